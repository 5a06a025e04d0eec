"""hyptokenizer_b200 -- B200 (sm_100a) native merge-loop hot path of HypTokenizer.

Public surface mirrors the reference's modules:
    hyptokenizer_b200.embedding.lorentz_model          <- embedding/lorentz_model.py
    hyptokenizer_b200.tokenizer.hyperbolic_merge       <- tokenizer/hyperbolic_merge.py
    hyptokenizer_b200.tokenizer.fast_hyperbolic_merge  <- tokenizer/fast_hyperbolic_merge.py
    hyptokenizer_b200.tokenizer.frequency_aware_hyperbolic_merge, .hierarchical_hyperbolic_merge,
    .compression_aware_tokenizer, .adaptive_curvature_tokenizer, .enhanced_fast_hyperbolic_merge
Everything computes through libhyptok_b200.so (include/hyptok_b200.h); there is no CPU path.
"""
__version__ = "0.1.0"
