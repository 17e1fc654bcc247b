// zpq_method.cpp -- SURVEY.md 8(a) row C1: model selection feeding the hot path.
//
// method string (+ block payload) -> block header bytes, for the context-model
// family only (no LZ77/BWT/E8E9 pre-processing, no PCOMP).  Restates
//   * libzpaq::compressBlock's expansion of digit methods (libzpaq.cpp:7556-7691),
//     including the data-dependent period analysis of level 5 (7655-7688),
//   * makeConfig's context-model section (libzpaq.cpp:7325-7534),
//   * the byte encoding the Compiler gives that ZPAQL source (libzpaq.cpp:2570-2690).
// Instead of generating ZPAQL text and tokenising it again, the COMP and HCOMP
// bytes are emitted directly by a small structured assembler.  Output is pinned
// byte for byte by tests/golden/headers.json (made with the unmodified reference).
#include "zpq_host.h"
#include <ctype.h>
#include <string.h>

namespace zpq {
namespace {

int lg(unsigned x) {   // floor(log2(x))+1, 0 for 0   (libzpaq.cpp:6566)
  int r = 0;
  while (x) { ++r; x >>= 1; }
  return r;
}
int popcount(unsigned x) { int r = 0; for (; x; x >>= 1) r += x & 1; return r; }

// ZPAQL opcodes used by the generator (indices of opcodelist, libzpaq.cpp:2456-2491)
enum {
  A_INC = 1, A_ZERO = 4, A_R = 7, B_INC = 9, C_DEC = 18, D_INC = 25, PD_ZERO = 52, JT = 39, JF = 47, R_A = 55,
  HALT = 56, HASH = 59, HASHD = 60, JMP = 63,
  A_C = 66, A_PB = 68, A_PC = 69, A_PD = 70, A_N = 71, B_A = 72, B_C = 74, D_A = 88, D_N = 95,
  PC_A = 104, PD_A = 112, PD_C = 114,
  ADD_B = 129, ADD_PB = 132, ADD_PC = 133, ADD_N = 135, SUB_C = 138, SUB_N = 143, MUL_N = 151, MOD_N = 167,
  AND_N = 175, SHL_N = 207, SHR_N = 215, EQ_N = 223, LT_N = 231, GT_N = 239
};

struct Asm {
  std::vector<uint8_t> b;
  std::vector<size_t> ifs;
  bool ok = true;
  void op(int o) { b.push_back((uint8_t)o); }
  void op(int o, int n) { if (n < 0 || n > 255) ok = false; b.push_back((uint8_t)o); b.push_back((uint8_t)n); }
  // IF / ELSE / ENDIF as the Compiler encodes them (libzpaq.cpp:2580-2622): short forward jumps
  void if_() { b.push_back(JF); b.push_back(0); ifs.push_back(b.size() - 1); }
  void else_() {
    if (ifs.empty()) { ok = false; return; }
    const size_t a = ifs.back(); ifs.pop_back();
    const size_t j = b.size() - a + 1;
    if (j > 127) ok = false;
    b[a] = (uint8_t)j;
    b.push_back(JMP); b.push_back(0); ifs.push_back(b.size() - 1);
  }
  void endif_() {
    if (ifs.empty()) { ok = false; return; }
    const size_t a = ifs.back(); ifs.pop_back();
    const size_t j = b.size() - a - 1;
    if (j > 127) ok = false;
    b[a] = (uint8_t)j;
  }
};

struct Comp {
  std::vector<uint8_t> b;
  bool ok = true;
  void add(int type, int a1 = -1, int a2 = -1, int a3 = -1, int a4 = -1, int a5 = -1) {
    b.push_back((uint8_t)type);
    const int a[5] = {a1, a2, a3, a4, a5};
    static const int len[10] = {0, 2, 3, 2, 3, 4, 6, 6, 3, 5};
    for (int i = 0; i < len[type] - 1; ++i) { if (a[i] < 0 || a[i] > 255) ok = false; b.push_back((uint8_t)a[i]); }
  }
};

std::string itos(long long x) { char s[32]; snprintf(s, sizeof(s), "%lld", x); return s; }

// compressBlock's expansion of "LB,R,t" digit methods (libzpaq.cpp:7556-7691).
int expand_digit_method(const std::string& m, const uint8_t* p, uint64_t n64, std::string* out, std::string* err) {
  const unsigned n = (unsigned)n64;
  const int arg0 = std::max(lg(n + 4095) - 20, 0);
  unsigned type = 0;
  {
    int commas = 0, arg[4] = {0, 0, 0, 0};
    for (size_t i = 1; i < m.size() && commas < 4; ++i) {
      if (m[i] == ',' || m[i] == '.') ++commas;
      else if (isdigit((unsigned char)m[i])) arg[commas] = arg[commas] * 10 + m[i] - '0';
    }
    type = commas == 0 ? 512u : (unsigned)(arg[1] * 4 + arg[2]);
  }
  const int level = m[0] - '0';
  const int doe8 = (type & 2) * 2;
  std::string method = "x" + itos(arg0);
  if (level <= 3) { *err = "levels 0-3 use LZ77/BWT pre-processing, which is outside the context-mixing path"; return ZPQ_E_METHOD; }
  if (level == 4) {
    if (type < 48 || type >= 900) { *err = "level 4 selects store/LZ77/BWT for this block type: outside the context-mixing path"; return ZPQ_E_METHOD; }
    method += "," + itos(doe8) + "ci1,1,1,1,2a";
    if (type & 1) method += "w";
    method += "m";
  }
  else {
    method += "," + itos(doe8);
    if (type & 1) method += "w2c0,1010,255i1";
    else method += "w1i1";
    method += "c256ci1,1,1,1,1,1,2a";
    // analyse the data: histogram of distances between equal bytes (libzpaq.cpp:7655-7666)
    const int NR = 1 << 12;
    std::vector<int> pt(256, 0), r(NR, 0);
    for (unsigned i = 0; i < n; ++i) {
      const int k = (int)i - pt[p[i]];
      if (k > 0 && k < NR) ++r[k];
      pt[p[i]] = (int)i;
    }
    int n1 = (int)n - r[1] - r[2] - r[3];
    for (int i = 0; i < 2; ++i) {
      int period = 0;
      double score = 0;
      int t = 0;
      for (int j = 5; j < NR && t < n1; ++j) {
        const double s = r[j] / (256.0 + n1 - t);
        if (s > score) { score = s; period = j; }
        t += r[j];
      }
      if (period > 4 && score > 0.1) {
        method += "c0,0," + itos(999 + period) + ",255i1";
        if (period <= 255) method += "c0," + itos(period) + "i1";
        n1 -= r[period];
        r[period] = 0;
      }
      else break;
    }
    method += "c0,2,0,255i1c0,3,0,0,255i1c0,4,0,0,0,255i1mm16ts19t0";
  }
  if (doe8) { *err = "E8E9 pre-processing (exe block type) is outside the context-mixing path"; return ZPQ_E_METHOD; }
  *out = method;
  return ZPQ_OK;
}

}  // namespace

int header_from_method(const char* method_, const uint8_t* data, uint64_t n, std::vector<uint8_t>* hdr, std::string* err) {
  std::string e0;
  if (!err) err = &e0;
  if (!method_ || !method_[0]) { *err = "empty method"; return ZPQ_E_METHOD; }
  if (n > 0x7fffffffull - 4096) { *err = "block too big"; return ZPQ_E_ARG; }
  std::string ms = method_;
  if (isdigit((unsigned char)ms[0])) {
    std::string x;
    const int rc = expand_digit_method(ms, data, n, &x, err);
    if (rc != ZPQ_OK) return rc;
    ms = x;
  }
  if (ms[0] != 'x') { *err = "Unsupported method"; return ZPQ_E_METHOD; }

  // ---- makeConfig: "{x}N1,N2...N9" into args (libzpaq.cpp:6893-6913) ----
  const char* m = ms.c_str();
  int args[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  ++m;
  for (int i = 0; i < 9 && (isdigit((unsigned char)*m) || *m == ',' || *m == '.');) {
    if (isdigit((unsigned char)*m)) args[i] = args[i] * 10 + *m - '0';
    else if (++i < 9) args[i] = 0;
    ++m;
  }
  if (args[1] != 0) { *err = "pre-processing (LZ77/BWT/E8E9) is outside the context-mixing path"; return ZPQ_E_METHOD; }
  if ((0x100000ull << args[0]) - 4096 < n && args[0] < 11) { /* the reference only asserts this */ }
  if (args[0] > 11) { *err = "block size argument too big"; return ZPQ_E_METHOD; }

  // ---- context model section (libzpaq.cpp:7325-7534) ----
  int ncomp = 0;
  const int membits = args[0] + 20;
  int sb = 5;
  Comp comp;
  Asm h;
  h.op(C_DEC); h.op(PC_A); h.op(ADD_N, 255); h.op(D_A); h.op(PD_C);   // c-- *c=a a+= 255 d=a *d=c

  while (*m && ncomp < 254) {
    std::vector<int> v;
    v.push_back((unsigned char)*m++);
    if (isdigit((unsigned char)*m)) {
      v.push_back(*m++ - '0');
      while (isdigit((unsigned char)*m) || *m == ',' || *m == '.') {
        if (isdigit((unsigned char)*m)) v.back() = v.back() * 10 + *m++ - '0';
        else { v.push_back(0); ++m; }
      }
    }

    if (v[0] == 'c') {   // context model: ICM or CM with optional special/masked contexts
      while (v.size() < 3) v.push_back(0);
      sb = 11;
      if (v[2] < 256) sb += lg((unsigned)v[2]); else sb += 6;
      for (size_t i = 3; i < v.size(); ++i) if (v[i] < 512) sb += popcount((unsigned)v[i]) * 3 / 4;
      if (sb > membits) sb = membits;
      if (v[1] % 1000 == 0) comp.add(ZT_ICM, sb - 6 - v[1] / 1000);
      else comp.add(ZT_CM, sb - 2 - v[1] / 1000, v[1] % 1000 - 1);
      h.op(D_N, ncomp); h.op(PD_ZERO);
      if (v[2] > 1 && v[2] <= 255) {   // periodic context
        h.op(A_C);
        if (lg((unsigned)v[2]) != lg((unsigned)v[2] - 1)) h.op(AND_N, v[2] - 1);
        else h.op(MOD_N, v[2]);
        h.op(HASHD);
      }
      else if (v[2] >= 1000 && v[2] <= 1255) {   // distance to last occurrence
        h.op(A_N, 255); h.op(ADD_N, v[2] - 1000); h.op(D_A); h.op(A_PD); h.op(SUB_C);
        h.op(GT_N, 255); h.if_(); h.op(A_N, 255); h.endif_();
        h.op(D_N, ncomp); h.op(HASHD);
      }
      for (size_t i = 3; i < v.size(); ++i) {   // masked context bytes
        if (i == 3) h.op(B_C);
        if (v[i] == 255) { h.op(A_PB); h.op(HASHD); }
        else if (v[i] > 0 && v[i] < 255) { h.op(A_PB); h.op(AND_N, v[i]); h.op(HASHD); }
        else if (v[i] >= 256 && v[i] < 512) {   // lz77 parse state or masked literal
          h.op(A_R, 1); h.op(GT_N, 1); h.if_();
            h.op(A_R, 2); h.op(LT_N, 64); h.if_();
              h.op(A_PB);
              if (v[i] < 511) h.op(AND_N, v[i] - 256);
              h.op(HASHD);
            h.else_();
              h.op(SHR_N, 6); h.op(HASHD); h.op(A_R, 1); h.op(HASHD);
            h.endif_();
          h.else_();
            h.op(A_N, 255); h.op(HASHD); h.op(A_R, 2); h.op(HASHD);
          h.endif_();
        }
        else if (v[i] >= 1256) {
          h.op(A_N, ((v[i] - 1000) >> 8) & 255); h.op(SHL_N, 8); h.op(ADD_N, (v[i] - 1000) & 255);
          h.op(ADD_B); h.op(B_A);
        }
        else if (v[i] > 1000) { h.op(A_N, v[i] - 1000); h.op(ADD_B); h.op(B_A); }
        if (v[i] < 512 && i < v.size() - 1) h.op(B_INC);
      }
      ++ncomp;
    }

    if (strchr("mts", v[0]) && ncomp > (v[0] == 't')) {   // MIX, MIX2, SSE
      if (v.size() <= 1) v.push_back(8);
      if (v.size() <= 2) v.push_back(24 + 8 * (v[0] == 's'));
      if (v[0] == 's' && v.size() <= 3) v.push_back(255);
      sb = 5 + v[1] * 3 / 4;
      if (v[0] == 'm') comp.add(ZT_MIX, v[1], 0, ncomp, v[2], 255);
      else if (v[0] == 't') comp.add(ZT_MIX2, v[1], ncomp - 1, ncomp - 2, v[2], 255);
      else comp.add(ZT_SSE, v[1], ncomp - 1, v[2], v[3]);
      if (v[1] > 8) {
        h.op(D_N, ncomp); h.op(PD_ZERO); h.op(B_C); h.op(A_ZERO);
        for (; v[1] >= 16; v[1] -= 8) {
          h.op(SHL_N, 8); h.op(ADD_PB);
          if (v[1] > 16) h.op(B_INC);
        }
        if (v[1] > 8) { h.op(SHL_N, 8); h.op(ADD_PB); h.op(SHR_N, 16 - v[1]); }
        h.op(SHL_N, 8); h.op(PD_A);
      }
      ++ncomp;
    }

    if (v[0] == 'i' && ncomp > 0) {   // ISSE chain
      h.op(D_N, ncomp - 1); h.op(B_C); h.op(A_PD); h.op(D_INC);
      for (size_t i = 1; i < v.size() && ncomp < 254; ++i) {
        for (int j = 0; j < v[i] % 10; ++j) {
          h.op(HASH);
          if (i < v.size() - 1 || j < v[i] % 10 - 1) h.op(B_INC);
          sb += 6;
        }
        h.op(PD_A);
        if (i < v.size() - 1) h.op(D_INC);
        if (sb > membits) sb = membits;
        comp.add(ZT_ISSE, sb - 6 - v[i] / 10, ncomp - 1);
        ++ncomp;
      }
    }

    if (v[0] == 'a') {   // MATCH
      if (v.size() <= 1) v.push_back(24);
      while (v.size() < 4) v.push_back(0);
      comp.add(ZT_MATCH, membits - v[3] - 2, membits - v[2]);
      h.op(D_N, ncomp); h.op(A_PD); h.op(MUL_N, v[1]); h.op(ADD_PC); h.op(A_INC); h.op(PD_A);
      sb = 5 + (membits - v[2]) * 3 / 4;
      ++ncomp;
    }

    if (v[0] == 'w') {   // word-model ICM-ISSE chain
      if (v.size() <= 1) v.push_back(1);
      if (v.size() <= 2) v.push_back(65);
      if (v.size() <= 3) v.push_back(26);
      if (v.size() <= 4) v.push_back(223);
      if (v.size() <= 5) v.push_back(20);
      if (v.size() <= 6) v.push_back(0);
      comp.add(ZT_ICM, membits - 6 - v[6]);
      for (int i = 1; i < v[1]; ++i) comp.add(ZT_ISSE, membits - 6 - v[6], ncomp + i - 1);
      h.op(A_PC); h.op(AND_N, v[4]); h.op(SUB_N, v[2]); h.op(AND_N, 255); h.op(LT_N, v[3]); h.if_();
      for (int i = 0; i < v[1]; ++i) {
        if (i == 0) h.op(D_N, ncomp); else h.op(D_INC);
        h.op(A_PD); h.op(MUL_N, v[5]); h.op(ADD_PC); h.op(A_INC); h.op(PD_A);
      }
      h.else_();
      for (int i = v[1] - 1; i > 0; --i) { h.op(D_N, ncomp + i - 1); h.op(A_PD); h.op(D_INC); h.op(PD_A); }
      h.op(D_N, ncomp); h.op(PD_ZERO);
      h.endif_();
      ncomp += v[1] - 1;
      sb = membits - v[6];
      ++ncomp;
    }
  }
  h.op(HALT);
  if (!h.ok || !comp.ok || !h.ifs.empty() || ncomp > 255) { *err = "method arguments out of range"; return ZPQ_E_METHOD; }

  // header: hsize[2] hh hm ph pm n COMP 0 HCOMP 0   ("comp 9 16 0 0", libzpaq.cpp:7307)
  const size_t hsize = 5 + comp.b.size() + 1 + h.b.size() + 1;
  if (hsize > 65535) { *err = "program too big"; return ZPQ_E_METHOD; }
  hdr->clear();
  hdr->push_back((uint8_t)(hsize & 255)); hdr->push_back((uint8_t)(hsize >> 8));
  hdr->push_back(9); hdr->push_back(16); hdr->push_back(0); hdr->push_back(0);
  hdr->push_back((uint8_t)ncomp);
  hdr->insert(hdr->end(), comp.b.begin(), comp.b.end());
  hdr->push_back(0);
  hdr->insert(hdr->end(), h.b.begin(), h.b.end());
  hdr->push_back(0);
  return ZPQ_OK;
}

}  // namespace zpq
