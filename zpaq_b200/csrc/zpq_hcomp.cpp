// zpq_hcomp.cpp -- turns a block's HCOMP bytecode into lane-parallel "recipes".
//
// The HCOMP programs that libzpaq's makeConfig writes (libzpaq.cpp:7325-7534)
// compute every context hash H[i] as a short fold over recent input bytes
// (a = (a + byte + 512) * 773 and relatives), sometimes seeded from the hash
// of the previous component.  Interpreting that bytecode with one warp costs
// ~134 ZPAQL instructions per input byte at level 5.  The
// recogniser below parses the bytecode back into one op list per component so
// that lane i evaluates H[i] itself, all lanes in parallel.
//
// Safety: the parse is followed by a self-check on the host -- the recipes and
// a straightforward interpreter of the same bytecode (ZPAQL::run semantics,
// libzpaq.cpp:1027-1262) are run side by side over a pseudo-random byte stream
// and must agree on every H[i] after every byte.  Any program that is not
// recognised or fails the check simply keeps rcp.ok = 0 and is interpreted on
// the device (vm_run), so foreign or hand-written models stay exact.
#include "zpq_host.h"
#include <string.h>
#include <new>

namespace zpq {
namespace {

enum {  // opcodes (opcodelist indices, libzpaq.cpp:2456-2491)
  A_INC = 1, A_ZERO = 4, B_INC = 9, C_DEC = 18, D_INC = 25, PD_ZERO = 52, JF = 47, HALT = 56, HASH = 59, HASHD = 60,
  JMP = 63, A_C = 66, A_PB = 68, A_PC = 69, A_PD = 70, A_N = 71, B_A = 72, B_C = 74, D_A = 88, D_N = 95,
  PC_A = 104, PD_A = 112, PD_C = 114, ADD_B = 129, ADD_PB = 132, ADD_PC = 133, ADD_N = 135, SUB_C = 138,
  SUB_N = 143, MUL_N = 151, MOD_N = 167, AND_N = 175, SHL_N = 207, SHR_N = 215, LT_N = 231, GT_N = 239
};

// ---- reference-semantics interpreter (host), used only for the self-check ----
struct HostVm {
  std::vector<uint8_t> m; std::vector<uint32_t> h; uint32_t r[256];
  uint32_t a = 0, b = 0, c = 0, d = 0; int f = 0;
  const uint8_t* code; int len;
  std::vector<uint8_t>* out = nullptr;      // OUT (opcode 57) appends here; PCOMP only
  long long max_steps = 1 << 20;            // per run(); the self-checks keep it small, post-processing raises it
  HostVm(int hh, int hm, const uint8_t* prog, int n) : m((size_t)1 << hm, 0), h((size_t)1 << hh, 0), code(prog), len(n) { memset(r, 0, sizeof(r)); }
  uint8_t fetch(int pc) const { return (pc >= 0 && pc < len) ? code[pc] : 0; }
  bool run(uint32_t input) {
    const uint32_t mm = (uint32_t)m.size() - 1, hm = (uint32_t)h.size() - 1;
    a = input;
    int pc = 0;
    for (long long guard = 0; guard < max_steps; ++guard) {
      const int op = fetch(pc++);
      if (op == 56) return true;
      if (op < 56) {
        const int tgt = op >> 3, act = op & 7;
        if (act == 7) {
          const int n = fetch(pc);
          switch (tgt) {
            case 0: a = r[n]; ++pc; break; case 1: b = r[n]; ++pc; break; case 2: c = r[n]; ++pc; break;
            case 3: d = r[n]; ++pc; break;
            case 4: if (f) pc += ((n + 128) & 255) - 127; else ++pc; break;
            case 5: if (!f) pc += ((n + 128) & 255) - 127; else ++pc; break;
            default: r[n] = a; ++pc; break;
          }
          continue;
        }
        if (act > 4 || op == 0) return false;
        uint32_t v;
        switch (tgt) { case 0: v = a; break; case 1: v = b; break; case 2: v = c; break; case 3: v = d; break;
                       case 4: v = m[b & mm]; break; case 5: v = m[c & mm]; break; default: v = h[d & hm]; }
        if (act == 0) { const uint32_t o = a; if (tgt == 4 || tgt == 5) { a = (o & ~255u) | (v & 255); v = o & 255; } else { a = v; v = o; } }
        else if (act == 1) ++v; else if (act == 2) --v; else if (act == 3) v = ~v; else v = 0;
        switch (tgt) { case 0: a = v; break; case 1: b = v; break; case 2: c = v; break; case 3: d = v; break;
                       case 4: m[b & mm] = (uint8_t)v; break; case 5: m[c & mm] = (uint8_t)v; break; default: h[d & hm] = v; }
        continue;
      }
      if (op < 64) {
        if (op == 57) { if (out) out->push_back((uint8_t)a); continue; }
        if (op == 59) { a = (a + m[b & mm] + 512) * 773; continue; }
        if (op == 60) { h[d & hm] = (h[d & hm] + a + 512) * 773; continue; }
        if (op == 63) { const int n = fetch(pc); pc += ((n + 128) & 255) - 127; continue; }
        return false;
      }
      if (op == 255) { const int t = fetch(pc) + 256 * fetch(pc + 1); if (t >= len) return false; pc = t; continue; }
      const int grp = op >> 3, src = op & 7;
      if (grp == 15 || grp >= 30) return false;
      uint32_t v;
      switch (src) { case 0: v = a; break; case 1: v = b; break; case 2: v = c; break; case 3: v = d; break;
                     case 4: v = m[b & mm]; break; case 5: v = m[c & mm]; break; case 6: v = h[d & hm]; break;
                     default: v = fetch(pc); ++pc; }
      switch (grp) {
        case 8: a = v; break; case 9: b = v; break; case 10: c = v; break; case 11: d = v; break;
        case 12: m[b & mm] = (uint8_t)v; break; case 13: m[c & mm] = (uint8_t)v; break; case 14: h[d & hm] = v; break;
        case 16: a += v; break; case 17: a -= v; break; case 18: a *= v; break;
        case 19: a = v ? a / v : 0; break; case 20: a = v ? a % v : 0; break;
        case 21: a &= v; break; case 22: a &= ~v; break; case 23: a |= v; break; case 24: a ^= v; break;
        case 25: a <<= (v & 31); break; case 26: a >>= (v & 31); break;
        case 27: f = (a == v); break; case 28: f = (a < v); break; default: f = (a > v); break;
      }
    }
    return false;
  }
};

// ---- host mirror of the device recipe evaluator (zpq_kernels2.cu: hcomp_recipes) ----
struct HostRcp {
  const ZRecipes& R; int n;
  std::vector<uint8_t> m; std::vector<uint32_t> lastpos; uint32_t hmask;
  uint32_t c = 0; uint32_t h[32];
  HostRcp(const ZRecipes& r, int n_, int hh, int hm) : R(r), n(n_), m((size_t)1 << hm, 0), lastpos((size_t)1 << hh, 0), hmask((1u << hh) - 1) { memset(h, 0, sizeof(h)); }
  void step(uint32_t byte) {
    const uint32_t mm = (uint32_t)m.size() - 1;
    --c; m[c & mm] = (uint8_t)byte;
    lastpos[(byte + 255) & hmask] = c;
    uint32_t old[32];
    memcpy(old, h, sizeof(old));
    for (uint32_t r = 0; r < R.nrounds; ++r)
      for (int l = 0; l < n; ++l) {
        if (R.round[l] != r) continue;
        uint32_t acc = R.start[l] == ZB_ZERO ? 0 : R.start[l] == ZB_OWN ? old[l] : h[R.base[l]];
        for (int t = 0; t < R.len[l]; ++t) {
          const uint32_t w = R.ops[t][l];
          const uint32_t op = w & 15, src = (w >> 4) & 15, arg = (w >> 8) & 255, off = w >> 16;
          uint32_t x = 0;
          if (src == ZS_BYTE) x = m[(c + off) & mm] & arg;
          else if (src == ZS_POSAND) x = c & arg;
          else if (src == ZS_POSMOD) x = arg ? c % arg : 0;
          else if (src == ZS_DIST) { x = lastpos[(255 + arg) & hmask] - c; if (x > 255) x = 255; }
          switch (op) {
            case ZR_FOLD: acc = (acc + x + 512) * 773; break;
            case ZR_SHL8ADD: acc = (acc << 8) + x; break;
            case ZR_SHR: acc >>= arg; break;
            case ZR_SHL: acc <<= arg; break;
            case ZR_ROLL: acc = acc * (w >> 16) + m[c & mm] + 1; break;
            case ZR_WORD: {
              const uint32_t w2 = R.ops[t + 1][l];
              const uint32_t v4 = arg, v2 = off & 255, v3 = off >> 8, mult = w2 & 255, first = (w2 >> 8) & 1;
              const uint32_t cur = m[c & mm];
              if ((((cur & v4) - v2) & 255) < v3) acc = acc * mult + cur + 1;
              else acc = first ? 0 : old[l - 1];
              ++t;
              break;
            }
            default: break;
          }
        }
        h[l] = acc;
      }
  }
};

struct Cursor {
  const uint8_t* p; int n, i;
  bool at(int k, int op) const { return i + k < n && p[i + k] == op; }
  int val(int k) const { return i + k < n ? p[i + k] : -1; }
};

struct LaneRcp { int start = ZB_ZERO, base = 0, round = 0; std::vector<uint32_t> ops; bool written = false; };

bool parse(const uint8_t* prog, int len, int ncomp, LaneRcp lanes[32], ZRecipes* R) {
  Cursor c{prog, len, 0};
  // prologue: c-- *c=a a+= 255 d=a *d=c
  if (!(c.at(0, C_DEC) && c.at(1, PC_A) && c.at(2, ADD_N) && c.val(3) == 255 && c.at(4, D_A) && c.at(5, PD_C))) return false;
  c.i += 6;
  auto lane_ok = [&](int l) { return l >= 0 && l < ncomp && l < 32; };
  while (c.i < len) {
    if (c.at(0, HALT)) break;
    // 'c' / mixer-context snippet: d= N *d=0 ...
    if (c.at(0, D_N) && c.at(2, PD_ZERO)) {
      const int l = c.val(1);
      if (!lane_ok(l)) return false;
      c.i += 3;
      LaneRcp& L = lanes[l];
      L = LaneRcp(); L.written = true;
      if (c.at(0, A_C) && c.at(1, AND_N) && c.at(3, HASHD)) { L.ops.push_back(ZR_WORD_OF(ZR_FOLD, ZS_POSAND, c.val(2), 0)); c.i += 4; }
      else if (c.at(0, A_C) && c.at(1, MOD_N) && c.at(3, HASHD)) { L.ops.push_back(ZR_WORD_OF(ZR_FOLD, ZS_POSMOD, c.val(2), 0)); c.i += 4; }
      else if (c.at(0, A_N) && c.val(1) == 255 && c.at(2, ADD_N) && c.at(4, D_A) && c.at(5, A_PD) && c.at(6, SUB_C) &&
               c.at(7, GT_N) && c.val(8) == 255 && c.at(9, JF) && c.val(10) == 2 && c.at(11, A_N) && c.val(12) == 255 &&
               c.at(13, D_N) && c.val(14) == l && c.at(15, HASHD)) {
        L.ops.push_back(ZR_WORD_OF(ZR_FOLD, ZS_DIST, c.val(3), 0));
        R->needs_lastpos = 1;
        c.i += 16;
      }
      if (c.at(0, B_C)) {
        ++c.i;
        int boff = 0;
        if (c.at(0, A_ZERO)) {   // byte assembly for mixer/SSE contexts
          ++c.i;
          if (!L.ops.empty()) return false;
          for (;;) {
            if (!(c.at(0, SHL_N) && c.val(1) == 8)) return false;
            if (c.at(2, PD_A)) { L.ops.push_back(ZR_WORD_OF(ZR_SHL, 0, 8, 0)); c.i += 3; break; }
            if (!c.at(2, ADD_PB)) return false;
            if (boff > 0xffff) return false;
            L.ops.push_back(ZR_WORD_OF(ZR_SHL8ADD, ZS_BYTE, 255, boff));
            c.i += 3;
            if (c.at(0, B_INC)) { ++boff; ++c.i; }
            else if (c.at(0, SHR_N)) { L.ops.push_back(ZR_WORD_OF(ZR_SHR, 0, c.val(1) & 31, 0)); c.i += 2; }
          }
        } else {
          for (;;) {
            if (boff > 0xffff) return false;
            if (c.at(0, A_PB) && c.at(1, HASHD)) { L.ops.push_back(ZR_WORD_OF(ZR_FOLD, ZS_BYTE, 255, boff)); c.i += 2; }
            else if (c.at(0, A_PB) && c.at(1, AND_N) && c.at(3, HASHD)) { L.ops.push_back(ZR_WORD_OF(ZR_FOLD, ZS_BYTE, c.val(2), boff)); c.i += 4; }
            else if (c.at(0, A_N) && c.at(2, SHL_N) && c.val(3) == 8 && c.at(4, ADD_N) && c.at(6, ADD_B) && c.at(7, B_A)) { boff += c.val(1) * 256 + c.val(5); c.i += 8; }
            else if (c.at(0, A_N) && c.at(2, ADD_B) && c.at(3, B_A)) { boff += c.val(1); c.i += 4; }
            else if (c.at(0, B_INC)) { ++boff; ++c.i; }
            else break;
          }
        }
      }
      continue;
    }
    // ISSE chain: d= N b=c a=*d d++ (hash [b++])* *d=a [d++] ...
    if (c.at(0, D_N) && c.at(2, B_C) && c.at(3, A_PD) && c.at(4, D_INC)) {
      const int base = c.val(1);
      if (!lane_ok(base)) return false;
      c.i += 5;
      LaneRcp cur;
      const bool pure = lanes[base].start == ZB_ZERO;
      if (pure) cur = lanes[base];                                  // pure function of recent bytes: inline it
      else { cur.start = ZB_LANE; cur.base = base; cur.round = lanes[base].round + 1; }
      int d = base + 1, boff = 0;
      for (;;) {
        while (c.at(0, HASH)) {
          if (boff > 0xffff) return false;
          cur.ops.push_back(ZR_WORD_OF(ZR_FOLD, ZS_BYTE, 255, boff));
          ++c.i;
          if (c.at(0, B_INC)) { ++boff; ++c.i; }
        }
        if (!c.at(0, PD_A)) return false;
        ++c.i;
        if (!lane_ok(d)) return false;
        lanes[d] = cur; lanes[d].written = true;
        if (c.at(0, D_INC)) {
          // a chain seeded by a stateful hash continues from the previous LINK (one pipeline stage back)
          if (!pure) { cur = LaneRcp(); cur.start = ZB_LANE; cur.base = d; cur.round = lanes[d].round + 1; }
          ++d; ++c.i; continue;
        }
        break;
      }
      continue;
    }
    // MATCH rolling hash: d= N a=*d a*= M a+=*c a++ *d=a
    if (c.at(0, D_N) && c.at(2, A_PD) && c.at(3, MUL_N) && c.at(5, ADD_PC) && c.at(6, A_INC) && c.at(7, PD_A)) {
      const int l = c.val(1);
      if (!lane_ok(l)) return false;
      LaneRcp& L = lanes[l];
      L = LaneRcp(); L.written = true; L.start = ZB_OWN;
      L.ops.push_back(ZR_WORD_OF(ZR_ROLL, 0, 0, c.val(4)));
      c.i += 8;
      continue;
    }
    // word model: a=*c a&= v4 a-= v2 a&= 255 a< v3 if ... else ... endif
    if (c.at(0, A_PC) && c.at(1, AND_N) && c.at(3, SUB_N) && c.at(5, AND_N) && c.val(6) == 255 && c.at(7, LT_N) && c.at(9, JF)) {
      const int v4 = c.val(2), v2 = c.val(4), v3 = c.val(8), jf = c.val(10);
      c.i += 11;
      const int then_start = c.i;
      if (!c.at(0, D_N)) return false;
      const int first = c.val(1);
      int cnt = 0, mult = -1;
      for (;;) {
        if (cnt == 0) { if (!(c.at(0, D_N) && c.val(1) == first)) return false; c.i += 2; }
        else { if (!c.at(0, D_INC)) break; ++c.i; }
        if (!(c.at(0, A_PD) && c.at(1, MUL_N) && c.at(3, ADD_PC) && c.at(4, A_INC) && c.at(5, PD_A))) return false;
        if (mult >= 0 && mult != c.val(2)) return false;
        mult = c.val(2);
        c.i += 6;
        ++cnt;
      }
      if (!c.at(0, JMP)) return false;
      const int jmp = c.val(1);
      c.i += 2;
      if (jf != c.i - then_start) return false;           // JF lands on the else branch
      const int else_start = c.i;
      for (int i = cnt - 1; i > 0; --i) {
        if (!(c.at(0, D_N) && c.val(1) == first + i - 1 && c.at(2, A_PD) && c.at(3, D_INC) && c.at(4, PD_A))) return false;
        c.i += 5;
      }
      if (!(c.at(0, D_N) && c.val(1) == first && c.at(2, PD_ZERO))) return false;
      c.i += 3;
      if (jmp != c.i - else_start) return false;          // JMP lands after the endif
      if (!lane_ok(first) || !lane_ok(first + cnt - 1) || v3 > 255) return false;
      for (int i = 0; i < cnt; ++i) {
        LaneRcp& L = lanes[first + i];
        L = LaneRcp(); L.written = true; L.start = ZB_OWN;
        L.ops.push_back(ZR_WORD_OF(ZR_WORD, 0, v4, v2 | v3 << 8));
        L.ops.push_back((uint32_t)mult | (uint32_t)(i == 0) << 8);
      }
      R->has_word = 1;
      continue;
    }
    return false;
  }
  if (!(c.at(0, HALT) && c.at(1, 0) && c.i + 2 == len)) return false;
  return true;
}

}  // namespace

void build_recipes(ZModel* M, const uint8_t* prog, int prog_len) {
  ZRecipes& R = M->rcp;
  memset(&R, 0, sizeof(R));
  if (M->hh != 9 || M->hm < 4 || M->hm > 24 || M->n > 32) return;   // makeConfig always writes "comp 9 16 0 0"
  LaneRcp lanes[32];
  ZRecipes tmp;
  memset(&tmp, 0, sizeof(tmp));
  if (!parse(prog, prog_len, (int)M->n, lanes, &tmp)) return;
  uint32_t nr = 1;
  for (uint32_t l = 0; l < M->n; ++l) {
    const LaneRcp& L = lanes[l];
    if (L.ops.size() > ZPQ_RCP_MAX) return;
    tmp.start[l] = (uint8_t)L.start; tmp.base[l] = (uint8_t)L.base; tmp.round[l] = (uint8_t)L.round;
    tmp.len[l] = (uint8_t)L.ops.size();
    for (size_t t = 0; t < L.ops.size(); ++t) tmp.ops[t][l] = L.ops[t];
    nr = std::max<uint32_t>(nr, (uint32_t)L.round + 1);
    if (L.start == ZB_LANE && (L.base >= (int)l)) return;   // bases precede their chain
  }
  tmp.nrounds = nr;
  // ---- self-check against the bytecode interpreter ----
  HostVm vm((int)M->hh, (int)M->hm, prog, prog_len);
  HostRcp rc(tmp, (int)M->n, (int)M->hh, (int)M->hm);
  uint32_t s = 12345;
  for (int i = 0; i < 6000; ++i) {
    s = s * 1664525u + 1013904223u;
    uint32_t byte = s >> 24;
    if ((i / 500) & 1) byte = "the quick\nbrown fox jumps over\n"[(s >> 24) % 31];   // text-like stretch: exercises the word model
    if (!vm.run(byte)) return;
    rc.step(byte);
    for (uint32_t l = 0; l < M->n; ++l)
      if (vm.h[l & ((1u << M->hh) - 1)] != rc.h[l]) return;
  }
  tmp.ok = 1;
  R = tmp;
}


// ---------------------------------------------------------------------------
// Flattened recipes for the wavefront encoder: component l's op list is prefixed with the lists of
// the components its hash is seeded from (ZB_LANE bases, earlier lanes of its word-model group), each
// as a segment with its own private state.  Verified against the per-lane recipes on a byte stream.
namespace {
struct FlatEval {
  uint32_t st[ZPQ_FLAT_SEGS];
  uint32_t dist_last; uint32_t c; uint64_t hist;
  std::vector<uint8_t> in;
  FlatEval() { memset(st, 0, sizeof(st)); dist_last = 0; c = 0; hist = 0; }
  uint32_t step(const uint32_t* ops, uint32_t len, uint32_t byte, uint32_t dist_byte) {
    --c; in.push_back((uint8_t)byte);
    const long long bpos = (long long)in.size() - 1;
    hist = (hist << 8) | byte;
    if (byte == dist_byte) dist_last = c;
    int k = -1; uint32_t acc = 0, prev_new = 0, prev_old = 0;
    for (uint32_t t = 0; t < len; ++t) {
      const uint32_t w = ops[t];
      const uint32_t op = w & 15, src = (w >> 4) & 15, arg = (w >> 8) & 255, off = w >> 16;
      if (op == ZR_SEG) {
        if (k >= 0) { prev_old = st[k]; st[k] = acc; prev_new = acc; }
        ++k;
        acc = src == ZB_ZERO ? 0 : (src == ZB_OWN ? st[k] : prev_new);
        continue;
      }
      uint32_t x = 0;
      if (src == ZS_BYTE) x = off < 8 ? ((uint32_t)(hist >> (8 * off)) & arg) : ((long long)off <= bpos ? (in[bpos - off] & arg) : 0);
      else if (src == ZS_POSAND) x = c & arg;
      else if (src == ZS_POSMOD) x = arg ? c % arg : 0;
      else if (src == ZS_DIST) { x = dist_last - c; if (x > 255) x = 255; }
      switch (op) {
        case ZR_FOLD: acc = (acc + x + 512) * 773; break;
        case ZR_SHL8ADD: acc = (acc << 8) + x; break;
        case ZR_SHR: acc >>= arg; break;
        case ZR_SHL: acc <<= arg; break;
        case ZR_ROLL: acc = acc * off + byte + 1; break;
        case ZR_WORD: {
          const uint32_t w2 = ops[t + 1];
          const uint32_t v2 = off & 255, v3 = off >> 8, mult = w2 & 255, first = (w2 >> 8) & 1;
          if ((((byte & arg) - v2) & 255) < v3) acc = acc * mult + byte + 1;
          else acc = first ? 0 : prev_old;
          ++t;
          break;
        }
        default: break;
      }
    }
    if (k >= 0) st[k] = acc;
    return acc;
  }
};
}  // namespace

void build_wave_plan(ZModel* M) {
  ZWave& W = M->wave;
  memset(&W, 0, sizeof(W));
  const ZRecipes& R = M->rcp;
  if (!R.ok || !M->pipe.ok || M->hm != 16) return;
  const int n = (int)M->n;
  for (int l = 0; l < n; ++l) {
    // chain of components whose state this hash needs, root first
    std::vector<int> chain;
    int cur = l;
    for (;;) {
      chain.insert(chain.begin(), cur);
      if (R.start[cur] == ZB_LANE) { cur = R.base[cur]; continue; }
      if (R.len[cur] >= 2 && (R.ops[0][cur] & 15u) == ZR_WORD && !((R.ops[1][cur] >> 8) & 1u) && cur > 0) { --cur; continue; }
      break;
    }
    if ((int)chain.size() > ZPQ_FLAT_SEGS) return;
    std::vector<uint32_t> f;
    int ndist = 0;
    for (size_t s = 0; s < chain.size(); ++s) {
      const int c = chain[s];
      f.push_back((uint32_t)ZR_SEG | (uint32_t)R.start[c] << 4);
      for (int t = 0; t < R.len[c]; ++t) {
        f.push_back(R.ops[t][c]);
        if ((R.ops[t][c] & 15u) == ZR_WORD) { f.push_back(R.ops[t + 1][c]); ++t; }
        else if (((R.ops[t][c] >> 4) & 15u) == ZS_DIST) ++ndist;
      }
    }
    if (f.size() > ZPQ_FLAT_MAX || ndist > 1) return;
    W.flat_len[l] = (uint32_t)f.size();
    for (size_t t = 0; t < f.size(); ++t) W.flat[l][t] = f[t];
  }
  // self-check against the per-lane recipes
  HostRcp rc(R, n, (int)M->hh, (int)M->hm);
  std::vector<FlatEval> fe(n);
  uint32_t s = 4242;
  for (int i = 0; i < 4000; ++i) {
    s = s * 1664525u + 1013904223u;
    uint32_t byte = s >> 24;
    if ((i / 400) & 1) byte = "the quick\nbrown fox jumps over\n"[(s >> 24) % 31];
    rc.step(byte);
    for (int l = 0; l < n; ++l)
      if (fe[l].step(W.flat[l], W.flat_len[l], byte, M->pipe.dist_byte[l]) != rc.h[l]) return;
  }
  // shape of the common recipes: order-n hash of the last n bytes (the ICM / ISSE chain of makeConfig), or constant 0
  for (int l = 0; l < n; ++l) {
    const uint32_t len = W.flat_len[l];
    bool fold = len >= 1 && len <= 9 && W.flat[l][0] == ((uint32_t)ZR_SEG | (uint32_t)ZB_ZERO << 4);
    for (uint32_t t = 1; t < len && fold; ++t)
      fold = W.flat[l][t] == ZR_WORD_OF(ZR_FOLD, ZS_BYTE, 255, t - 1);
    W.fold_n[l] = fold ? (uint8_t)len : 0;        // len = n + 1 (the segment op, then n folds)
  }
  W.dcoder = M->pipe.depth[n - 1] + 1;
  if (W.dcoder > 15) return;
  W.ok = 1;
}


// ---------------------------------------------------------------------------
// Slot-scheduled hashes (ZHashPlan, zpq_dev.h): merge the flat op lists of all lanes into one sequence of
// slots with a warp-uniform op kind.  Greedy: the next slot takes the kind that is the pending op of the most
// lanes (ties: the smaller ZH_* value), every lane whose pending op has that kind advances.  The result is
// checked against FlatEval on the same byte stream before it is marked usable.
namespace {
struct PlanEval {      // host mirror of hplan_eval (zpq_kcommon.cuh)
  uint32_t st[32][ZPQ_FLAT_SEGS];
  uint32_t dist_last[32]; uint32_t c; uint64_t hist;
  std::vector<uint8_t> in;
  PlanEval() { memset(st, 0, sizeof(st)); memset(dist_last, 0, sizeof(dist_last)); c = 0; hist = 0; }
  void step(const ZHashPlan& P, int n, uint32_t byte, const uint8_t* dist_byte, uint32_t h[32]) {
    --c; in.push_back((uint8_t)byte);
    const long long bpos = (long long)in.size() - 1;
    hist = (hist << 8) | byte;
    uint32_t acc[32] = {0}, prev_new[32] = {0}, prev_old[32] = {0};
    for (int l = 0; l < n; ++l) if (byte == dist_byte[l]) dist_last[l] = c;
    for (uint32_t s = 0; s < P.nslots; ++s)
      for (int l = 0; l < n; ++l) {
        if (!((P.slot[s].lanes >> l) & 1u)) continue;
        const uint32_t kind = P.slot[s].kind;
        const uint32_t w = P.w[s][l], src = (w >> 4) & 15, arg = (w >> 8) & 255, off = w >> 16;
        uint32_t x = 0;
        if (kind == ZH_FOLD_BYTE || kind == ZH_SHL8_BYTE) x = (uint32_t)(hist >> (8 * off)) & arg;
        else if (kind == ZH_FOLD_OTHER || kind == ZH_SHL8_OTHER) {
          if (src == ZS_BYTE) x = (long long)off <= bpos ? (in[bpos - off] & arg) : 0;
          else if (src == ZS_POSAND) x = c & arg;
          else if (src == ZS_POSMOD) x = arg ? c % arg : 0;
          else if (src == ZS_DIST) { x = dist_last[l] - c; if (x > 255) x = 255; }
        }
        switch (kind) {
          case ZH_SEG: {
            const uint32_t k = (w >> 8) & 3u;
            if (k > 0) { prev_old[l] = st[l][k - 1]; st[l][k - 1] = acc[l]; prev_new[l] = acc[l]; }
            acc[l] = src == ZB_ZERO ? 0 : (src == ZB_OWN ? st[l][k] : prev_new[l]);
            break;
          }
          case ZH_FOLD_BYTE: case ZH_FOLD_OTHER: acc[l] = (acc[l] + x + 512) * 773; break;
          case ZH_SHL8_BYTE: case ZH_SHL8_OTHER: acc[l] = (acc[l] << 8) + x; break;
          case ZH_SHR: acc[l] >>= arg; break;
          case ZH_SHL: acc[l] <<= arg; break;
          case ZH_ROLL: acc[l] = acc[l] * off + byte + 1; break;
          case ZH_WORD: {
            const uint32_t w2 = P.w2[s][l];
            const uint32_t v2 = off & 255, v3 = off >> 8, mult = w2 & 255, first = (w2 >> 8) & 1;
            if ((((byte & arg) - v2) & 255) < v3) acc[l] = acc[l] * mult + byte + 1;
            else acc[l] = first ? 0 : prev_old[l];
            break;
          }
          default: break;
        }
      }
    for (int l = 0; l < n; ++l)
      if (P.fold_n[l]) {
        uint32_t a = 0;
        for (uint32_t k = 0; k + 1 < P.fold_n[l]; ++k) a = (a + ((uint32_t)(hist >> (8 * k)) & 255u) + 512u) * 773u;
        acc[l] = a;
      }
    for (int l = 0; l < n; ++l) { st[l][P.last_seg[l]] = acc[l]; h[l] = acc[l]; }
  }
};
}  // namespace

void build_hash_plan(ZModel* M) {
  ZHashPlan& P = M->hplan;
  memset(&P, 0, sizeof(P));
  const ZWave& W = M->wave;
  if (!W.ok) return;
  const int n = (int)M->n;
  // per lane: the op list with ZR_WORD's second word folded into its op and segment indexes made explicit
  struct Op { uint32_t kind, w, w2; };
  std::vector<std::vector<Op>> ops(n);
  for (int l = 0; l < n; ++l) {
    if (W.fold_n[l]) { P.fold_n[l] = W.fold_n[l]; P.fold_lanes |= 1u << l; P.last_seg[l] = 0; continue; }   // folded outside the slots
    int k = -1;
    for (uint32_t t = 0; t < W.flat_len[l]; ++t) {
      const uint32_t w = W.flat[l][t], kind = w & 15u;
      if (kind == ZR_END) break;
      const uint32_t src = (w >> 4) & 15u, off = w >> 16;
      const bool reg_byte = src == ZS_BYTE && off < 8;        // one of the last eight bytes: in the history register
      if (kind == ZR_SEG) { ++k; ops[l].push_back(Op{ZH_SEG, (w & 0xffu) | (uint32_t)k << 8, 0}); }
      else if (kind == ZR_WORD) { ops[l].push_back(Op{ZH_WORD, w, W.flat[l][t + 1]}); ++t; }
      else if (kind == ZR_FOLD) ops[l].push_back(Op{reg_byte ? (uint32_t)ZH_FOLD_BYTE : (uint32_t)ZH_FOLD_OTHER, w, 0});
      else if (kind == ZR_SHL8ADD) ops[l].push_back(Op{reg_byte ? (uint32_t)ZH_SHL8_BYTE : (uint32_t)ZH_SHL8_OTHER, w, 0});
      else if (kind == ZR_SHR) ops[l].push_back(Op{ZH_SHR, w, 0});
      else if (kind == ZR_SHL) ops[l].push_back(Op{ZH_SHL, w, 0});
      else if (kind == ZR_ROLL) ops[l].push_back(Op{ZH_ROLL, w, 0});
      else return;                                  // an op this evaluator does not know
    }
    if (k < 0 || ops[l].empty() || ops[l][0].kind != ZH_SEG) return;
    P.last_seg[l] = (uint8_t)k;
  }
  std::vector<size_t> at(n, 0);
  uint32_t ns = 0;
  for (;;) {
    int count[16] = {0};
    bool any = false;
    for (int l = 0; l < n; ++l) if (at[l] < ops[l].size()) { ++count[ops[l][at[l]].kind]; any = true; }
    if (!any) break;
    int best = -1;
    for (int kd = 0; kd < 16; ++kd) if (count[kd] > 0 && (best < 0 || count[kd] > count[best])) best = kd;
    if (ns >= ZPQ_HSLOTS) return;
    P.slot[ns].kind = (uint32_t)best;
    for (int l = 0; l < n; ++l)
      if (at[l] < ops[l].size() && (int)ops[l][at[l]].kind == best) {
        P.slot[ns].lanes |= 1u << l; P.w[ns][l] = ops[l][at[l]].w; P.w2[ns][l] = ops[l][at[l]].w2; ++at[l];
      }
    ++ns;
  }
  P.nslots = ns;
  // per-role slot lists (two-warp decoders)
  {
    uint32_t role_mask[2] = {0, 0};
    for (int l = 0; l < n; ++l) {
      const uint32_t t = M->comp[l].type;
      const int role = (t == ZT_ICM || t == ZT_ISSE || t == ZT_MATCH) ? 0 : 1;        // the split zpq_kernels7.cu uses
      role_mask[role] |= 1u << l;
    }
    for (int r = 0; r < 2; ++r) {
      P.role_n[r] = 0; P.role_fold[r] = P.fold_lanes & role_mask[r];
      for (uint32_t s2 = 0; s2 < ns; ++s2) if (P.slot[s2].lanes & role_mask[r]) P.role_slots[r][P.role_n[r]++] = (uint8_t)s2;
    }
  }
  // self-check against the flat form (which build_wave_plan checked against the recipes, which build_recipes
  // checked against the bytecode)
  std::vector<FlatEval> fe(n);
  PlanEval pe;
  uint32_t s = 977;
  for (int i = 0; i < 4000; ++i) {
    s = s * 1664525u + 1013904223u;
    uint32_t byte = s >> 24;
    if ((i / 400) & 1) byte = "the quick\nbrown fox jumps over\n"[(s >> 24) % 31];
    uint32_t h[32];
    pe.step(P, n, byte, M->pipe.dist_byte, h);
    for (int l = 0; l < n; ++l)
      if (fe[l].step(W.flat[l], W.flat_len[l], byte, M->pipe.dist_byte[l]) != h[l]) return;
  }
  P.ok = 1;
}


// ---------------------------------------------------------------------------
// Post-processing of one block's model output (PostProcessor::write, libzpaq.cpp:2193-2241), host side: the first
// byte through the model is 0 (PASS: the rest is the payload) or 1 (PROG: two length bytes and a PCOMP program follow;
// every later byte, and 0xffffffff at the end of each segment, is run through that program, whose OUT instructions
// produce the payload).  This is not part of the GPU hot path: it runs on the host after the block is decoded.
int postprocess_block(int ph, int pm, const uint8_t* model_out, const uint64_t* seg_len, int nseg,
                      std::vector<uint8_t>* payload, std::vector<uint64_t>* seg_payload, std::string* err) {
  payload->clear(); seg_payload->assign((size_t)std::max(nseg, 0), 0);
  // ZPAQL::init (libzpaq.cpp:1009-1024) rejects these sizes ("H too big" / "M too big"); 1 << 255 is not a size
  if (ph < 0 || ph > 32 || pm < 0 || pm > 32) { if (err) *err = "H or M too big"; return ZPQ_E_COMPONENT; }
  int state = 0;                 // 0 first byte, 1 PASS, 2/3 length bytes, 4 program bytes, 5 running
  size_t want = 0;
  std::vector<uint8_t> prog;
  HostVm* vm = nullptr;
  int rc = ZPQ_OK;
  const uint8_t* p = model_out;
  for (int sg = 0; sg < nseg && rc == ZPQ_OK; ++sg) {
    const size_t before = payload->size();
    for (uint64_t i = 0; i < seg_len[sg] && rc == ZPQ_OK; ++i) {
      const uint32_t c = *p++;
      switch (state) {
        case 0:
          if (c > 1) { if (err) *err = "unknown post processing type"; rc = ZPQ_E_CORRUPT; }
          else state = c == 0 ? 1 : 2;
          break;
        case 1: payload->push_back((uint8_t)c); break;
        case 2: want = c; state = 3; break;
        case 3:
          want += (size_t)c * 256;
          if (want < 1) { if (err) *err = "Empty PCOMP"; rc = ZPQ_E_CORRUPT; }
          prog.reserve(want); state = 4;
          break;
        case 4:
          prog.push_back((uint8_t)c);
          if (prog.size() == want) {
            try { vm = new HostVm(ph, pm, prog.data(), (int)prog.size()); }     // ZPAQL::initp: H = 2^ph words, M = 2^pm bytes
            catch (const std::bad_alloc&) { if (err) *err = "Out of memory"; return ZPQ_E_NOMEM; }
            vm->out = payload; vm->max_steps = (long long)1 << 40;
            state = 5;
          }
          break;
        default:
          if (!vm->run(c)) { if (err) *err = "ZPAQL execution error"; rc = ZPQ_E_VM; }
      }
    }
    if (rc == ZPQ_OK) {
      if (state == 5) { if (!vm->run(0xffffffffu)) { if (err) *err = "ZPAQL execution error"; rc = ZPQ_E_VM; } }
      else if (state != 1) { if (err) *err = "Unexpected EOS"; rc = ZPQ_E_CORRUPT; }
    }
    (*seg_payload)[sg] = payload->size() - before;
  }
  delete vm;
  return rc;
}

}  // namespace zpq
