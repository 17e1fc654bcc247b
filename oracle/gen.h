/* oracle/gen.h -- deterministic synthetic block generators (TEST INFRASTRUCTURE).
 *
 * Shared by the reference driver (oracle/ref_driver.cpp), the C oracle
 * (oracle/zpq_oracle.c) and, through ctypes, tests/ and bench.py so that the
 * CPU reference and the GPU path always see byte-identical inputs.
 *
 * Generators follow BASELINE.md section 3 / SURVEY.md section 8(d):
 *   LCG  s = s*1664525 + 1013904223 (mod 2^32), one step per draw.
 *   kind 0 "text"   words from a fixed 16-word list joined by ' ' or '\n'
 *   kind 1 "mixed"  "etaoin shrdlu\n"[(s>>24) % 14]
 *   kind 2 "random" s>>24
 *   kind 3 "period" 24-byte records: 16 low-entropy bytes + 8 counter bytes
 *                   (fires the level-5 period detector, libzpaq.cpp:7655-7688)
 * Per-block seed = base seed + block index (caller adds).
 */
#ifndef ZPQ_ORACLE_GEN_H
#define ZPQ_ORACLE_GEN_H
#include <stdint.h>
#include <stddef.h>

static inline uint32_t zpq_lcg(uint32_t* s) {
  *s = *s * 1664525u + 1013904223u;
  return *s;
}

static inline void zpq_gen_block(int kind, uint32_t seed, unsigned char* out, size_t n) {
  static const char* const words[16] = {
    "the", "of", "and", "compression", "model", "context", "mixing", "bit",
    "predict", "archive", "block", "state", "table", "hash", "journal", "update"};
  uint32_t s = seed;
  size_t i = 0;
  switch (kind) {
    case 0:
      while (i < n) {
        const char* w = words[(zpq_lcg(&s) >> 24) & 15];
        while (*w && i < n) out[i++] = (unsigned char)*w++;
        if (i < n) out[i++] = ((zpq_lcg(&s) >> 24) & 7) == 0 ? '\n' : ' ';
      }
      break;
    case 1:
      for (; i < n; ++i) out[i] = (unsigned char)"etaoin shrdlu\n"[(zpq_lcg(&s) >> 24) % 14];
      break;
    case 2:
      for (; i < n; ++i) out[i] = (unsigned char)(zpq_lcg(&s) >> 24);
      break;
    default: {
      uint32_t rec = 0;
      while (i < n) {
        for (int k = 0; k < 16 && i < n; ++k) out[i++] = (unsigned char)('a' + ((zpq_lcg(&s) >> 24) & 3));
        for (int k = 0; k < 8 && i < n; ++k) out[i++] = (unsigned char)((rec >> (8 * (k & 3))) + k);
        ++rec;
      }
    }
  }
}
#endif
