"""CPU oracle for the HypTokenizer merge-loop hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``hyptokenizer_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` may.  It restates, op for op, the
arithmetic of the reference (pure Python + torch CPU eager fp32):

* ``oracle.lorentz``   <- embedding/lorentz_model.py
* ``oracle.merge``     <- tokenizer/hyperbolic_merge.py, fast_hyperbolic_merge.py,
                          frequency_aware_hyperbolic_merge.py
* ``oracle.sumorder``  <- the ATen CPU reduction orders those files inherit
                          (SURVEY.md Appendix D), as plain numpy
* ``oracle/pair_count.c`` <- frequency_aware_hyperbolic_merge.py:92-112 in C

Parity is PINNED: ``oracle/gen_golden.py`` imports the unmodified reference from
/root/reference (in the build container) and writes ``tests/golden/*.json``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors
bit for bit.
"""
