// zpq_method.cpp -- SURVEY.md 8(a) row C1: model selection feeding the hot path.
//
// method string (+ block payload) -> block header bytes, for the context-model family only (no LZ77 / BWT / E8E9
// pre-processing, no PCOMP).  The output has to equal what libzpaq::compressBlock -> makeConfig -> Compiler write for
// the same method (libzpaq.cpp:7543-7691, 7325-7534, 2570-2690) and is pinned byte for byte by tests/golden/headers.json
// (168 headers made with the unmodified reference) plus a fuzz test against the live reference on the GPU box.
//
// Structure (this file's own, not the reference's text-generating flow):
//   1. tokenize()        method text  -> tokens {letter, numeric arguments}
//   2. level_tokens()    digit methods ("4", "50", "5,R,T") -> the same token list a spelled-out method would give;
//                        level 5 asks dominant_gaps() for up to two periodic structures in the payload
//   3. family builders   one per token letter, looked up in kFamilies: each appends component records and HCOMP code to a
//                        Draft.  HCOMP code is assembled from opcode tables with parameter slots (Snip), not from ZPAQL text
//   4. serialise()       Draft -> header bytes
#include "zpq_host.h"
#include <ctype.h>
#include <string.h>

namespace zpq {
namespace {

// ---------------------------------------------------------------- small helpers
int bit_length(unsigned x) { int r = 0; for (; x; x >>= 1) ++r; return r; }           // 0 -> 0, 1 -> 1, 255 -> 8, 256 -> 9
int ones(unsigned x) { int r = 0; for (; x; x &= x - 1) ++r; return r; }

// ---------------------------------------------------------------- 1. tokens
struct Token {
  char letter = 0;
  std::vector<int> arg;
  int at(size_t i, int dflt) const { return i < arg.size() ? arg[i] : dflt; }
};

// A token is one character followed -- only if the very next character is a digit -- by numbers separated by ',' or '.'.
// A separator opens a new (zero) number, digits extend the last one; anything else ends the token.  Characters that are
// not family letters become tokens too and are ignored later, exactly as stray characters are in the reference's scan.
std::vector<Token> tokenize(const char* s) {
  std::vector<Token> out;
  while (*s) {
    Token t;
    t.letter = *s++;
    if (isdigit((unsigned char)*s)) {
      t.arg.push_back(0);
      for (; isdigit((unsigned char)*s) || *s == ',' || *s == '.'; ++s) {
        if (isdigit((unsigned char)*s)) t.arg.back() = t.arg.back() * 10 + (*s - '0');
        else t.arg.push_back(0);
      }
    }
    out.push_back(t);
  }
  return out;
}
Token tok(char letter, std::initializer_list<int> a = {}) { Token t; t.letter = letter; t.arg.assign(a.begin(), a.end()); return t; }

// ---------------------------------------------------------------- 2. digit methods
// Gap statistics of level 5 (libzpaq.cpp:7655-7688).  For every byte the distance to the previous byte of the same value
// is counted (1 .. 4095).  A "dominant gap" is the distance g >= 5 with the largest share hist[g] / (256 + pool - seen),
// where pool = bytes not explained by gaps 1..3 (nor by an earlier dominant gap) and seen = the counts of the gaps
// 5 .. g-1; the scan stops once seen reaches pool.  A gap is accepted when its share exceeds 0.1; at most two are taken,
// the second after removing the first from the statistics.
std::vector<int> dominant_gaps(const uint8_t* p, unsigned n) {
  enum { kGaps = 1 << 12 };
  std::vector<int> hist(kGaps, 0);
  int last_seen[256];
  memset(last_seen, 0, sizeof(last_seen));
  for (unsigned i = 0; i < n; ++i) {
    const int gap = (int)i - last_seen[p[i]];
    if (gap > 0 && gap < kGaps) ++hist[gap];
    last_seen[p[i]] = (int)i;
  }
  std::vector<int> found;
  int pool = (int)n - hist[1] - hist[2] - hist[3];
  while (found.size() < 2) {
    int best = 0, seen = 0;
    double best_share = 0;
    for (int g = 5; g < kGaps && seen < pool; ++g) {
      const double share = hist[g] / (256.0 + pool - seen);
      if (share > best_share) { best_share = share; best = g; }
      seen += hist[g];
    }
    if (best <= 4 || !(best_share > 0.1)) break;
    found.push_back(best);
    pool -= hist[best];
    hist[best] = 0;
  }
  return found;
}

// "LB[,R,T]": L = level digit, B = block size digits (ignored here: the block decides), R,T -> block type R*4+T
// (bit 0 text, bit 1 exe; 512 when no type is given).  Every branch of libzpaq.cpp:7596-7650: store (an empty token list),
// LZ77 (lz[] = its five numbers), BWT, context mixing, each with or without E8E9 (*prep = makeConfig's second number).
int level_tokens(const std::string& m, uint64_t n64, const std::vector<int>& gaps, int* arg0, int* prep, int lz[5], std::vector<Token>* out, std::string* err) {
  const unsigned n = (unsigned)n64;
  *arg0 = std::max(bit_length(n + 4095) - 20, 0);
  int field[4] = {0, 0, 0, 0}, nsep = 0;
  for (size_t i = 1; i < m.size() && nsep < 4; ++i) {
    if (m[i] == ',' || m[i] == '.') ++nsep;
    else if (isdigit((unsigned char)m[i])) field[nsep] = field[nsep] * 10 + (m[i] - '0');
  }
  const unsigned type = nsep == 0 ? 512u : (unsigned)(field[1] * 4 + field[2]);
  const bool text = (type & 1u) != 0, exe = (type & 2u) != 0;
  const int level = m[0] - '0';
  const int e8 = exe ? 4 : 0;                  // E8E9 address rewriting in front of whatever the level selects
  std::vector<Token>& t = *out;
  t.clear();
  *prep = e8;
  // LZ77 parameters (makeConfig's numbers 3..7: minimum match, second context order, log2 bucket, table bits, look-ahead);
  // `tbl` = hash table bits for this block size, `sa` = "use the suffix array" (libzpaq.cpp:7586-7587)
  const int tbl = 19 + *arg0 + (*arg0 <= 6 ? 1 : 0), sa = 21 + *arg0;
  auto lz77 = [&](int kind, int a2, int a3, int a4, int a5, int a6) { *prep = kind + e8; lz[0] = a2; lz[1] = a3; lz[2] = a4; lz[3] = a5; lz[4] = a6; };
  if (level == 0) { *prep = 0; return ZPQ_OK; }                      // "0B": stored, no model (an empty token list)
  if (level == 1) {
    if (type < 40) { *prep = 0; return ZPQ_OK; }                     // hard to compress: store
    if (type < 80) lz77(1, 4, 0, 1, 15, 0);
    else if (type < 128) lz77(1, 4, 0, 2, 16, 0);
    else if (type < 256) lz77(1, 4, 0, 2, tbl, 0);
    else if (type < 960) lz77(1, 5, 0, 3, tbl, 0);
    else lz77(1, 6, 0, 3, tbl, 0);
    return ZPQ_OK;
  }
  if (level == 2) {
    if (type < 32) { *prep = 0; return ZPQ_OK; }
    if (type < 64) lz77(1, 4, 0, 3, tbl, 0);
    else lz77(1, 4, 0, 7, sa, 1);
    return ZPQ_OK;
  }
  if (level == 3) {
    if (type < 20) { *prep = 0; return ZPQ_OK; }
    if (type < 48) { lz77(1, 4, 0, 3, tbl, 0); return ZPQ_OK; }
    if (type >= 640 || text) {                                       // BWT for text and for highly redundant blocks
      *prep = 3 + e8;
      t.push_back(tok('c')); t.push_back(tok('i', {1}));
      return ZPQ_OK;
    }
    lz77(2, 12, 0, 7, sa, 1);                                        // byte-aligned LZ77, orders 0-1 over literals and parse state
    t.push_back(tok('c', {0, 0, 511})); t.push_back(tok('i', {2}));
    return ZPQ_OK;
  }
  if (level == 4) {
    if (type < 12) { *prep = 0; return ZPQ_OK; }
    if (type < 24) { lz77(1, 4, 0, 3, tbl, 0); return ZPQ_OK; }
    if (type < 48) {
      // the reference glues "1c0,0,511" to the table-size number (libzpaq.cpp:7637): the sixth number becomes e.g. 211,
      // which still means "suffix array", and there is no look-ahead number
      lz77(2, 5, 0, 7, sa * 10 + 1, 0);
      t.push_back(tok('c', {0, 0, 511}));
      return ZPQ_OK;
    }
    if (type >= 900) {
      *prep = 3 + e8;
      t.push_back(tok('c')); t.push_back(tok('i', {1}));
      return ZPQ_OK;
    }
    // order 0..5 chain, match model, optional word model, one mixer
    t.push_back(tok('c'));
    t.push_back(tok('i', {1, 1, 1, 1, 2}));
    t.push_back(tok('a'));
    if (text) t.push_back(tok('w'));
    t.push_back(tok('m'));
    return ZPQ_OK;
  }
  // level 5: word model (two links plus a line-position context for text), order-0 CM, order 0..8 chain, match model,
  // up to two periodic models, three sparse contexts, two mixers, a final MIX2 / SSE / MIX2 stage
  if (text) { t.push_back(tok('w', {2})); t.push_back(tok('c', {0, 1010, 255})); t.push_back(tok('i', {1})); }
  else { t.push_back(tok('w', {1})); t.push_back(tok('i', {1})); }
  t.push_back(tok('c', {256}));
  t.push_back(tok('c'));
  t.push_back(tok('i', {1, 1, 1, 1, 1, 1, 2}));
  t.push_back(tok('a'));
  for (int gap : gaps) {
    t.push_back(tok('c', {0, 0, 999 + gap, 255})); t.push_back(tok('i', {1}));            // the byte one period back
    if (gap <= 255) { t.push_back(tok('c', {0, gap})); t.push_back(tok('i', {1})); }      // position inside the period
  }
  t.push_back(tok('c', {0, 2, 0, 255})); t.push_back(tok('i', {1}));
  t.push_back(tok('c', {0, 3, 0, 0, 255})); t.push_back(tok('i', {1}));
  t.push_back(tok('c', {0, 4, 0, 0, 0, 255})); t.push_back(tok('i', {1}));
  t.push_back(tok('m')); t.push_back(tok('m', {16})); t.push_back(tok('t')); t.push_back(tok('s', {19})); t.push_back(tok('t', {0}));
  return ZPQ_OK;
}

// ---------------------------------------------------------------- 3. draft model and HCOMP assembly
// ZPAQL opcodes used by the generator (indexes of the opcode list, libzpaq.cpp:2456-2491)
enum {
  A_INC = 1, A_DEC = 2, A_ZERO = 4, A_R = 7, B_INC = 9, C_DEC = 18, D_INC = 25, PD_ZERO = 52, JT = 39, JF = 47, R_A = 55,
  HALT = 56, HASH = 59, HASHD = 60, JMP = 63,
  A_C = 66, A_PB = 68, A_PC = 69, A_PD = 70, A_N = 71, B_A = 72, B_C = 74, D_A = 88, D_N = 95,
  PC_A = 104, PD_A = 112, PD_C = 114,
  ADD_B = 129, ADD_PB = 132, ADD_PC = 133, ADD_N = 135, SUB_C = 138, SUB_N = 143, MUL_N = 151, MOD_N = 167,
  AND_N = 175, SHL_N = 207, SHR_N = 215, EQ_N = 223, LT_N = 231, GT_N = 239
};

// One instruction of a snippet: opcode, then NONE, a literal operand 0..255, or P(k) = the k-th parameter of the call.
// IF / ELSE / ENDIF are pseudo-ops resolved to the short forward jumps the Compiler writes (libzpaq.cpp:2580-2622).
enum { NONE = -1, IF_ = 1000, ELSE_ = 1001, ENDIF_ = 1002 };
constexpr int P(int k) { return 256 + k; }
struct Snip { int op; int operand; };

struct Record { int type; int a[5]; };          // one COMP entry: type byte and up to five arguments

struct Draft {
  std::vector<Record> comps;
  std::vector<uint8_t> code;
  std::vector<size_t> open_jumps;               // operand positions of JF / JMP waiting for their target
  bool ok = true;
  int membits = 20;                             // log2 of the block size class
  int sb = 5;                                   // running "size bits" hint handed from one family to the next

  int count() const { return (int)comps.size(); }
  void component(int type, int a1 = -1, int a2 = -1, int a3 = -1, int a4 = -1, int a5 = -1) {
    static const int kBytes[10] = {0, 2, 3, 2, 3, 4, 6, 6, 3, 5};
    Record r{type, {a1, a2, a3, a4, a5}};
    for (int i = 0; i + 1 < kBytes[type]; ++i) if (r.a[i] < 0 || r.a[i] > 255) ok = false;
    comps.push_back(r);
  }
  void byte(int v) { if (v < 0 || v > 255) ok = false; code.push_back((uint8_t)v); }
  void close_jump(size_t extra) {
    if (open_jumps.empty()) { ok = false; return; }
    const size_t at = open_jumps.back(); open_jumps.pop_back();
    const size_t dist = code.size() - at - 1 + extra;
    if (dist > 127) ok = false;
    code[at] = (uint8_t)dist;
  }
  template <size_t N> void emit(const Snip (&s)[N], std::initializer_list<int> params = {}) {
    for (const Snip& i : s) {
      if (i.op == IF_) { code.push_back(JF); code.push_back(0); open_jumps.push_back(code.size() - 1); continue; }
      if (i.op == ELSE_) { close_jump(2); code.push_back(JMP); code.push_back(0); open_jumps.push_back(code.size() - 1); continue; }
      if (i.op == ENDIF_) { close_jump(0); continue; }
      code.push_back((uint8_t)i.op);
      if (i.operand == NONE) continue;
      if (i.operand < 256) byte(i.operand);
      else {
        const size_t k = (size_t)(i.operand - 256);
        byte(k < params.size() ? *(params.begin() + k) : -1);
      }
    }
  }
};

// every program starts by pushing the new byte into the M ring and remembering where that byte value was last seen
const Snip kProlog[] = {{C_DEC, NONE}, {PC_A, NONE}, {ADD_N, 255}, {D_A, NONE}, {PD_C, NONE}};
// byte-aligned LZ77: R1 = 1 + bytes until the next code, R2 = first byte of the current code (libzpaq.cpp:7339-7357);
// p0 = the length of the post-processor program that precedes the codes in the first segment
const Snip kParseState[] = {{A_R, 1}, {EQ_N, 0}, {IF_, NONE}, {A_N, P(0)}, {ELSE_, NONE}, {EQ_N, 1}, {IF_, NONE}, {A_PC, NONE}, {R_A, 2},
                            {GT_N, 63}, {IF_, NONE}, {SHR_N, 6}, {A_INC, NONE}, {A_INC, NONE}, {ELSE_, NONE}, {A_INC, NONE}, {A_INC, NONE}, {ENDIF_, NONE},
                            {ELSE_, NONE}, {A_DEC, NONE}, {ENDIF_, NONE}, {ENDIF_, NONE}, {R_A, 1}};
const Snip kClearH[] = {{D_N, P(0)}, {PD_ZERO, NONE}};                                   // H[p0] = 0
const Snip kPosAnd[] = {{A_C, NONE}, {AND_N, P(0)}, {HASHD, NONE}};                      // position & (2^k - 1)
const Snip kPosMod[] = {{A_C, NONE}, {MOD_N, P(0)}, {HASHD, NONE}};                      // position % period
const Snip kLastSeen[] = {{A_N, 255}, {ADD_N, P(0)}, {D_A, NONE}, {A_PD, NONE}, {SUB_C, NONE}, {GT_N, 255}, {IF_, NONE},
                          {A_N, 255}, {ENDIF_, NONE}, {D_N, P(1)}, {HASHD, NONE}};      // min(255, distance to the last byte p0)
const Snip kByteFull[] = {{A_PB, NONE}, {HASHD, NONE}};
const Snip kByteMasked[] = {{A_PB, NONE}, {AND_N, P(0)}, {HASHD, NONE}};
// LZ77 parse-state context (R1 = state, R2 = literal-or-length byte); p0 = literal mask
const Snip kParseMasked[] = {{A_R, 1}, {GT_N, 1}, {IF_, NONE}, {A_R, 2}, {LT_N, 64}, {IF_, NONE}, {A_PB, NONE}, {AND_N, P(0)}, {HASHD, NONE},
                             {ELSE_, NONE}, {SHR_N, 6}, {HASHD, NONE}, {A_R, 1}, {HASHD, NONE}, {ENDIF_, NONE},
                             {ELSE_, NONE}, {A_N, 255}, {HASHD, NONE}, {A_R, 2}, {HASHD, NONE}, {ENDIF_, NONE}};
const Snip kParseFull[] = {{A_R, 1}, {GT_N, 1}, {IF_, NONE}, {A_R, 2}, {LT_N, 64}, {IF_, NONE}, {A_PB, NONE}, {HASHD, NONE},
                           {ELSE_, NONE}, {SHR_N, 6}, {HASHD, NONE}, {A_R, 1}, {HASHD, NONE}, {ENDIF_, NONE},
                           {ELSE_, NONE}, {A_N, 255}, {HASHD, NONE}, {A_R, 2}, {HASHD, NONE}, {ENDIF_, NONE}};
const Snip kSkipFar[] = {{A_N, P(0)}, {SHL_N, 8}, {ADD_N, P(1)}, {ADD_B, NONE}, {B_A, NONE}};      // b += p0 * 256 + p1
const Snip kSkipNear[] = {{A_N, P(0)}, {ADD_B, NONE}, {B_A, NONE}};                                // b += p0
const Snip kPackOpen[] = {{D_N, P(0)}, {PD_ZERO, NONE}, {B_C, NONE}, {A_ZERO, NONE}};
const Snip kPackByte[] = {{SHL_N, 8}, {ADD_PB, NONE}};
const Snip kPackPart[] = {{SHL_N, 8}, {ADD_PB, NONE}, {SHR_N, P(0)}};
const Snip kPackClose[] = {{SHL_N, 8}, {PD_A, NONE}};
const Snip kChainOpen[] = {{D_N, P(0)}, {B_C, NONE}, {A_PD, NONE}, {D_INC, NONE}};                 // a = H[p0], d = p0 + 1, b = newest byte
const Snip kRoll[] = {{A_PD, NONE}, {MUL_N, P(0)}, {ADD_PC, NONE}, {A_INC, NONE}, {PD_A, NONE}};   // H[d] = H[d] * p0 + byte + 1
const Snip kWordTest[] = {{A_PC, NONE}, {AND_N, P(0)}, {SUB_N, P(1)}, {AND_N, 255}, {LT_N, P(2)}, {IF_, NONE}};
const Snip kShiftDown[] = {{D_N, P(0)}, {A_PD, NONE}, {D_INC, NONE}, {PD_A, NONE}};                // H[p0 + 1] = H[p0]

void one(Draft& d, int op) { d.code.push_back((uint8_t)op); }
void one(Draft& d, int op, int n) { d.code.push_back((uint8_t)op); d.byte(n); }

// 'c': an ICM (count limit 0) or CM whose context is built from an optional positional term and masked recent bytes.
//   args: [limit*? + slot bits], [special], [byte selectors ...]
//   special 2..255: position modulo it; 1000..1255: distance to the last occurrence of byte (special - 1000)
//   selector 255 / 1..254: that byte, whole or masked; 0: skip it; 256..511: LZ77 parse state (mask = selector - 256);
//   1001..: skip (selector - 1000) bytes
void family_context(Draft& d, const Token& t) {
  const int spec0 = t.at(0, 0), special = t.at(1, 0);
  int bits = 11 + (special < 256 ? bit_length((unsigned)special) : 6);
  for (size_t i = 2; i < t.arg.size(); ++i) if (t.arg[i] < 512) bits += ones((unsigned)t.arg[i]) * 3 / 4;
  d.sb = std::min(bits, d.membits);
  const int me = d.count();
  if (spec0 % 1000 == 0) d.component(ZT_ICM, d.sb - 6 - spec0 / 1000);
  else d.component(ZT_CM, d.sb - 2 - spec0 / 1000, spec0 % 1000 - 1);
  d.emit(kClearH, {me});
  if (special > 1 && special <= 255) {
    const bool pow2 = bit_length((unsigned)special) != bit_length((unsigned)special - 1);
    if (pow2) d.emit(kPosAnd, {special - 1}); else d.emit(kPosMod, {special});
  }
  else if (special >= 1000 && special <= 1255) d.emit(kLastSeen, {special - 1000, me});
  for (size_t i = 2; i < t.arg.size(); ++i) {
    const int sel = t.arg[i];
    if (i == 2) one(d, B_C);
    if (sel == 255) d.emit(kByteFull);
    else if (sel > 0 && sel < 255) d.emit(kByteMasked, {sel});
    else if (sel >= 256 && sel < 511) d.emit(kParseMasked, {sel - 256});
    else if (sel == 511) d.emit(kParseFull);
    else if (sel >= 1256) d.emit(kSkipFar, {((sel - 1000) >> 8) & 255, (sel - 1000) & 255});
    else if (sel > 1000) d.emit(kSkipNear, {sel - 1000});
    if (sel < 512 && i + 1 < t.arg.size()) one(d, B_INC);
  }
}

// 'm' / 't' / 's': MIX over everything so far, MIX2 / SSE over the last component(s).  A context wider than 8 bits packs
// whole recent bytes (and one partial byte) above the c8 byte.
void family_final(Draft& d, const Token& t) {
  const int have = d.count();
  if (have <= (t.letter == 't' ? 1 : 0)) return;
  int width = t.at(0, 8);
  const int rate = t.at(1, t.letter == 's' ? 32 : 24);
  d.sb = 5 + width * 3 / 4;
  if (t.letter == 'm') d.component(ZT_MIX, width, 0, have, rate, 255);
  else if (t.letter == 't') d.component(ZT_MIX2, width, have - 1, have - 2, rate, 255);
  else d.component(ZT_SSE, width, have - 1, rate, t.at(2, 255));
  if (width <= 8) return;
  d.emit(kPackOpen, {have});
  for (; width >= 16; width -= 8) {
    d.emit(kPackByte);
    if (width > 16) one(d, B_INC);
  }
  if (width > 8) d.emit(kPackPart, {16 - width});
  d.emit(kPackClose);
}

// 'i': a chain of ISSE links, each hashing (arg % 10) more bytes on top of the previous link's context
void family_chain(Draft& d, const Token& t) {
  if (d.count() == 0) return;
  d.emit(kChainOpen, {d.count() - 1});
  for (size_t i = 0; i < t.arg.size() && d.count() < 254; ++i) {
    const int more = t.arg[i] % 10;
    const bool last_link = i + 1 == t.arg.size();
    for (int j = 0; j < more; ++j) {
      one(d, HASH);
      if (!last_link || j + 1 < more) one(d, B_INC);
      d.sb += 6;
    }
    one(d, PD_A);
    if (!last_link) one(d, D_INC);
    d.sb = std::min(d.sb, d.membits);
    d.component(ZT_ISSE, d.sb - 6 - t.arg[i] / 10, d.count() - 1);
  }
}

// 'a': MATCH with a rolling hash of multiplier arg0
void family_match(Draft& d, const Token& t) {
  const int mult = t.at(0, 24), buf_less = t.at(1, 0), idx_less = t.at(2, 0);
  const int me = d.count();
  d.component(ZT_MATCH, d.membits - idx_less - 2, d.membits - buf_less);
  one(d, D_N, me);
  d.emit(kRoll, {mult});
  d.sb = 5 + (d.membits - buf_less) * 3 / 4;
}

// 'w': word model -- an ICM and (links - 1) ISSE over rolling hashes of the current and previous words; a byte b is a
// letter when ((b & mask) - first) mod 256 < span
void family_word(Draft& d, const Token& t) {
  const int links = t.at(0, 1), first = t.at(1, 65), span = t.at(2, 26), mask = t.at(3, 223), mult = t.at(4, 20), less = t.at(5, 0);
  const int me = d.count();
  d.component(ZT_ICM, d.membits - 6 - less);
  for (int i = 1; i < links; ++i) d.component(ZT_ISSE, d.membits - 6 - less, me + i - 1);
  d.emit(kWordTest, {mask, first, span});
  for (int i = 0; i < links; ++i) {                      // letter: extend every word hash
    if (i == 0) one(d, D_N, me); else one(d, D_INC);
    d.emit(kRoll, {mult});
  }
  const Snip kElse[] = {{ELSE_, NONE}};
  d.emit(kElse);
  for (int i = links - 1; i > 0; --i) d.emit(kShiftDown, {me + i - 1});     // word ends: hashes move one link down
  d.emit(kClearH, {me});
  const Snip kEnd[] = {{ENDIF_, NONE}};
  d.emit(kEnd);
  d.sb = d.membits - less;
}

struct Family { char letter; void (*build)(Draft&, const Token&); };
const Family kFamilies[] = {
  {'c', family_context}, {'m', family_final}, {'t', family_final}, {'s', family_final},
  {'i', family_chain}, {'a', family_match}, {'w', family_word},
};

// ---------------------------------------------------------------- 4. header bytes
// hsize[2] hh hm ph pm n COMP 0 HCOMP 0   ("comp 9 16 0 0", libzpaq.cpp:7307; BWT blocks: ph = pm = 20 + size class, 7137)
int serialise(const Draft& d, int ph, int pm, std::vector<uint8_t>* hdr, std::string* err) {
  static const int kBytes[10] = {0, 2, 3, 2, 3, 4, 6, 6, 3, 5};
  std::vector<uint8_t> comp;
  for (const Record& r : d.comps) {
    comp.push_back((uint8_t)r.type);
    for (int i = 0; i + 1 < kBytes[r.type]; ++i) comp.push_back((uint8_t)r.a[i]);
  }
  const size_t hsize = 5 + comp.size() + 1 + d.code.size() + 1;
  if (hsize > 65535) { *err = "program too big"; return ZPQ_E_METHOD; }
  hdr->clear();
  const uint8_t fixed[7] = {(uint8_t)(hsize & 255), (uint8_t)(hsize >> 8), 9, 16, (uint8_t)ph, (uint8_t)pm, (uint8_t)d.count()};
  hdr->insert(hdr->end(), fixed, fixed + 7);
  hdr->insert(hdr->end(), comp.begin(), comp.end());
  hdr->push_back(0);
  hdr->insert(hdr->end(), d.code.begin(), d.code.end());
  hdr->push_back(0);
  return ZPQ_OK;
}

}  // namespace

// a digit method of level 5 and up looks at the payload (gap statistics); everything else is decided by the text alone
bool method_needs_gaps(const char* method) {
  return method && isdigit((unsigned char)method[0]) && method[0] >= '5';
}

void payload_gaps(const uint8_t* data, uint64_t n, int gaps[2]) {
  gaps[0] = gaps[1] = 0;
  const std::vector<int> g = dominant_gaps(data, (unsigned)n);
  for (size_t i = 0; i < g.size() && i < 2; ++i) gaps[i] = g[i];
}

int header_from_method(const char* method_, const uint8_t* data, uint64_t n, std::vector<uint8_t>* hdr, std::string* err) {
  int gaps[2] = {0, 0};
  if (method_needs_gaps(method_) && n <= 0x7fffffffull - 4096) payload_gaps(data, n, gaps);
  return header_from_stats(method_, n, gaps, hdr, err);
}

// method (+ payload statistics) -> header bytes and the pre-processing the payload needs before it enters the model:
// *prep = 0 none, 3 BWT, 4 E8E9, 7 E8E9 then BWT (makeConfig's second number, libzpaq.cpp:6894); LZ77 kinds are not built
int block_plan(const char* method_, uint64_t n, const int gaps_in[2], std::vector<uint8_t>* hdr, int* prep, int* size_class_out, int lz_out[5], std::string* err) {
  std::string e0;
  if (!err) err = &e0;
  if (!method_ || !method_[0]) { *err = "empty method"; return ZPQ_E_METHOD; }
  if (n > 0x7fffffffull - 4096) { *err = "block too big"; return ZPQ_E_ARG; }
  const std::string ms = method_;
  std::vector<Token> tokens;
  int size_class = 0, preprocessing = 0;
  int lz[5] = {0, 0, 0, 0, 0};
  if (isdigit((unsigned char)ms[0])) {
    std::vector<int> gaps;
    for (int i = 0; i < 2 && gaps_in && gaps_in[i] > 0; ++i) gaps.push_back(gaps_in[i]);
    const int rc = level_tokens(ms, n, gaps, &size_class, &preprocessing, lz, &tokens, err);
    if (rc != ZPQ_OK) return rc;
  }
  else if (ms[0] == 'x') {
    // "x" N1 [, N2 ...]: block size class, pre-processing type, further arguments of the pre-processor;
    // up to nine numbers, then the component tokens (libzpaq.cpp:6893-6913)
    const char* s = ms.c_str() + 1;
    int head[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, k = 0;
    for (; k < 9 && (isdigit((unsigned char)*s) || *s == ',' || *s == '.'); ++s) {
      if (isdigit((unsigned char)*s)) head[k] = head[k] * 10 + (*s - '0');
      else if (++k < 9) head[k] = 0;
    }
    size_class = head[0]; preprocessing = head[1];
    for (int i = 0; i < 5; ++i) lz[i] = head[2 + i];
    tokens = tokenize(s);
  }
  else { *err = "Unsupported method"; return ZPQ_E_METHOD; }
  if (isdigit((unsigned char)ms[0]) && ms[0] == '0') {
    // level 0: "comp 0 0 0 0 0 hcomp end" -- no machines, no components, no program (libzpaq.cpp:6914-6916)
    static const uint8_t kNone[9] = {7, 0, 0, 0, 0, 0, 0, 0, 0};
    hdr->assign(kNone, kNone + 9);
    if (prep) *prep = 0;
    if (size_class_out) *size_class_out = size_class;
    if (lz_out) for (int i = 0; i < 5; ++i) lz_out[i] = 0;
    return ZPQ_OK;
  }
  if (preprocessing < 0 || preprocessing > 7) { *err = "Unsupported method"; return ZPQ_E_METHOD; }
  const int lzkind = preprocessing & 3;
  if ((lzkind == 1 && lz[0] < 4) || (lzkind == 2 && lz[0] < 1)) { *err = "match length $3 too small"; return ZPQ_E_METHOD; }
  if ((lzkind == 1 || lzkind == 2) && (lz[0] > 255 || lz[2] > 16 || lz[3] < 1 || lz[3] > 1000)) { *err = "LZ77 arguments out of range"; return ZPQ_E_METHOD; }
  if (size_class > 11) { *err = "block size argument too big"; return ZPQ_E_METHOD; }
  // the inverse BWT keeps the block (n + 5 bytes) in a 2^(20 + size class) byte array
  if ((preprocessing & 3) == 3 && n + 4096 > ((uint64_t)0x100000 << size_class)) { *err = "block does not fit the method's block size"; return ZPQ_E_METHOD; }

  Draft d;
  d.membits = size_class + 20;
  d.emit(kProlog);
  if (lzkind == 2) d.emit(kParseState, {111 + (preprocessing >= 4 ? 57 : 0)});
  for (const Token& t : tokens) {
    if (d.count() >= 254) break;
    for (const Family& f : kFamilies) if (f.letter == t.letter) { f.build(d, t); break; }
  }
  one(d, HALT);
  if (!d.ok || !d.open_jumps.empty() || d.count() > 255) { *err = "method arguments out of range"; return ZPQ_E_METHOD; }
  const bool bwt = lzkind == 3;
  if (prep) *prep = preprocessing;
  if (size_class_out) *size_class_out = size_class;
  if (lz_out) for (int i = 0; i < 5; ++i) lz_out[i] = lz[i];
  // the post-processor's machine: BWT keeps list and bytes (H and M of 2^(20 + size class)), LZ77 only the bytes (libzpaq.cpp:6923, 7076, 7137)
  return serialise(d, bwt ? size_class + 20 : 0, (bwt || lzkind == 1 || lzkind == 2) ? size_class + 20 : 0, hdr, err);
}

// header only, for callers that feed the model themselves (zpq_model_from_method, zpq_header_from_stats): a method that
// transforms the payload is refused here -- the header alone would describe a block the caller cannot produce
int header_from_stats(const char* method_, uint64_t n, const int gaps_in[2], std::vector<uint8_t>* hdr, std::string* err) {
  std::string e0;
  if (!err) err = &e0;
  int prep = 0;
  const int rc = block_plan(method_, n, gaps_in, hdr, &prep, nullptr, nullptr, err);
  if (rc != ZPQ_OK) return rc;
  if (prep != 0 || (hdr->size() > 6 && (*hdr)[6] == 0)) {
    hdr->clear(); *err = "this method stores or pre-processes the payload (E8E9 / BWT / LZ77): use zpq_compress_blocks"; return ZPQ_E_METHOD;
  }
  return ZPQ_OK;
}

}  // namespace zpq
