/* pair_count.c -- CPU restatement (plain C) of the reference's pair counting.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows
 * tokenizer/frequency_aware_hyperbolic_merge.py:92-112 with `_merge_rules == {}`
 * (SURVEY.md 3.5): for each line of the text-mode file (universal newlines: LF, CR, CRLF),
 * strip() leading/trailing str.isspace() code points, then count adjacent code-point pairs.
 * Written line by line, the way the reference iterates, NOT the way the GPU kernel does it.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int is_space_cp(uint32_t c) {
  if (c < 0x80) return (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x20);
  return c == 0x85 || c == 0xa0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200a) || c == 0x2028 ||
         c == 0x2029 || c == 0x202f || c == 0x205f || c == 0x3000;
}

typedef struct { uint64_t key, val; } slot_t;
typedef struct { slot_t *s; uint64_t cap, used; } map_t;

static void map_init(map_t *m, uint64_t cap) {
  m->cap = cap; m->used = 0;
  m->s = (slot_t *)malloc(cap * sizeof(slot_t));
  for (uint64_t k = 0; k < cap; ++k) { m->s[k].key = UINT64_MAX; m->s[k].val = 0; }
}
static void map_add(map_t *m, uint64_t key, uint64_t inc);
static void map_grow(map_t *m) {
  map_t b; map_init(&b, m->cap * 2);
  for (uint64_t k = 0; k < m->cap; ++k) if (m->s[k].key != UINT64_MAX) map_add(&b, m->s[k].key, m->s[k].val);
  free(m->s); *m = b;
}
static void map_add(map_t *m, uint64_t key, uint64_t inc) {
  if (m->used * 2 > m->cap) map_grow(m);
  uint64_t h = (key * 0x9E3779B97F4A7C15ULL) >> 20;
  for (;;) {
    h &= m->cap - 1;
    if (m->s[h].key == key) { m->s[h].val += inc; return; }
    if (m->s[h].key == UINT64_MAX) { m->s[h].key = key; m->s[h].val = inc; m->used++; return; }
    ++h;
  }
}

/* decode one line [p, e) into code points */
static size_t decode_line(const uint8_t *p, const uint8_t *e, uint32_t *out) {
  size_t n = 0;
  while (p < e) {
    uint32_t b = *p, c; int need;
    if (b < 0x80) { c = b; need = 1; }
    else if (b >= 0xF0) { c = b & 7; need = 4; }
    else if (b >= 0xE0) { c = b & 15; need = 3; }
    else { c = b & 31; need = 2; }
    for (int k = 1; k < need && p + k < e; ++k) c = (c << 6) | (p[k] & 0x3F);
    p += need;
    out[n++] = c;
  }
  return n;
}

/* Returns the number of distinct non-ASCII pairs written to (keys, vals), or -1 if more than
 * max_other.  ascii_counts[128*128]. */
long pair_count_oracle(const uint8_t *text, size_t n, uint64_t *ascii_counts, uint64_t *keys,
                       uint64_t *vals, long max_other) {
  memset(ascii_counts, 0, 128 * 128 * sizeof(uint64_t));
  map_t m; map_init(&m, 1024);
  size_t cap = 1 << 16;
  uint32_t *cps = (uint32_t *)malloc(cap * sizeof(uint32_t));
  size_t pos = 0;
  while (pos < n) {
    size_t end = pos;
    while (end < n && text[end] != '\n' && text[end] != '\r') ++end;
    if (end - pos + 1 > cap) { cap = (end - pos + 1) * 2; cps = (uint32_t *)realloc(cps, cap * sizeof(uint32_t)); }
    size_t len = decode_line(text + pos, text + end, cps);
    size_t a = 0, b = len;
    while (a < b && is_space_cp(cps[a])) ++a;        /* line.strip() */
    while (b > a && is_space_cp(cps[b - 1])) --b;
    for (size_t k = a; k + 1 < b; ++k) {
      uint32_t x = cps[k], y = cps[k + 1];
      if (x < 128 && y < 128) ascii_counts[x * 128 + y]++;
      else map_add(&m, ((uint64_t)x << 32) | y, 1);
    }
    pos = end + 1;
  }
  long w = 0;
  for (uint64_t k = 0; k < m.cap; ++k)
    if (m.s[k].key != UINT64_MAX) {
      if (w >= max_other) { w = -1; break; }
      keys[w] = m.s[k].key; vals[w] = m.s[k].val; ++w;
    }
  free(m.s); free(cps);
  return w;
}
