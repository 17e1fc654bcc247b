// zpq_model.cpp -- host side "model compiler": block header bytes -> ZModel
// (arena layout, init ops, predict step list, guarded HCOMP image), plus the
// model-independent tables.
//
// Reference behaviour restated: ZPAQL::read (libzpaq.cpp:887-930) for header
// structure, Predictor::init (libzpaq.cpp:1723-1851) for argument validation
// (same error strings) and table sizes, ZPAQL::memory (986-1006) for sizes.
#include "zpq_host.h"
#include "zpq_kernels.h"
#include <math.h>
#include <string.h>

namespace zpq {

static const int kCompLen[10] = {0, 2, 3, 2, 3, 4, 6, 6, 3, 5};  // libzpaq.cpp:714

// ---------------------------------------------------------------- tables
// Bit-history state machine of the ZPAQ specification; libzpaq ships the result
// as 1024 bytes of data (libzpaq.cpp:726-855).  Pinned by tests/golden/tables.json.
namespace {
int ns_count(int n0, int n1) {
  static const int bound[6] = {20, 48, 15, 8, 6, 5};
  if (n0 < n1) std::swap(n0, n1);
  if (n0 < 0 || n1 < 0 || n1 >= 6 || n0 > bound[n1]) return 0;
  return 1 + (n1 > 0 && n0 + n1 <= 17);
}
int ns_discount(int n) { return (n >= 1) + (n >= 2) + (n >= 3) + (n >= 4) + (n >= 5) + (n >= 7) + (n >= 8); }
void ns_step(int& n0, int& n1, int y) {
  if (n0 < n1) { ns_step(n1, n0, 1 - y); return; }
  if (y) { ++n1; n0 = ns_discount(n0); } else { ++n0; n1 = ns_discount(n1); }
  while (!ns_count(n0, n1)) {
    if (n1 < 2) --n0;
    else { n0 = (n0 * (n1 - 1) + (n1 / 2)) / n1; --n1; }
  }
}
}  // namespace

void build_state_table(uint8_t ns[1024]) {
  const int N = 50;
  static uint8_t t[50][50][2];
  memset(t, 0, sizeof(t));
  int state = 0;
  for (int i = 0; i < N; ++i)
    for (int n1 = 0; n1 <= i; ++n1) {
      const int n0 = i - n1, k = ns_count(n0, n1);
      if (k) { t[n0][n1][0] = (uint8_t)state; t[n0][n1][1] = (uint8_t)(state + k - 1); state += k; }
    }
  memset(ns, 0, 1024);
  for (int n0 = 0; n0 < N; ++n0)
    for (int n1 = 0; n1 < N; ++n1)
      for (int y = 0; y < ns_count(n0, n1); ++y) {
        const int s = t[n0][n1][y];
        int a = n0, b = n1;
        ns_step(a, b, 0); ns[s * 4 + 0] = t[a][b][0];
        a = n0; b = n1;
        ns_step(a, b, 1); ns[s * 4 + 1] = t[a][b][1];
        ns[s * 4 + 2] = (uint8_t)n0; ns[s * 4 + 3] = (uint8_t)n1;
      }
}

void build_tables(ZTables* T) {
  // squash/stretch closed forms: libzpaq.cpp:1736,1743 (comments there)
  static int16_t stretch[32768];
  for (int i = 0; i < 4096; ++i) {
    int q = (int)(32768.0 / (1 + exp((i - 2048) * (-1.0 / 64))));
    if (i < 1376) q = 0;
    if (i >= 2720) q = 32767;
    T->squash[i] = (uint16_t)q;
  }
  for (int i = 0; i < 32768; ++i)
    stretch[i] = (int16_t)((int)(log((i + 0.5) / (32767.5 - i)) * 64 + 0.5 + 100000) - 100000);
  memcpy(T->stretch_hi, stretch + 16384, sizeof(T->stretch_hi));
  T->dt2k[0] = 0;
  for (int i = 1; i < 256; ++i) T->dt2k[i] = (uint16_t)(2048 / i);
  for (int i = 0; i < 1024; ++i) T->dt[i] = (1 << 17) / (i * 2 + 3) * 2;
  uint8_t ns[1024];
  build_state_table(ns);
  for (int s = 0; s < 256; ++s) {
    T->ns[s] = ns[s * 4] | ns[s * 4 + 1] << 8 | ns[s * 4 + 2] << 16 | (uint32_t)ns[s * 4 + 3] << 24;
    const int n0 = ns[s * 4 + 2], n1 = ns[s * 4 + 3];
    const int cminit = ((n1 * 2 + 1) << 22) / (n0 + n1 + 1);       // StateTable::cminit, libzpaq.h:1106
    T->icm_init[s] = (uint32_t)cminit;
    int w = stretch[cminit >> 8] * 1024;                           // libzpaq.cpp:1832
    if (w < -(1 << 19)) w = -(1 << 19);
    if (w >= (1 << 19)) w = (1 << 19) - 1;
    T->isse_init[s * 2] = 1 << 15;
    T->isse_init[s * 2 + 1] = (uint32_t)w;
  }
}

// ---------------------------------------------------------------- header -> model
namespace {
struct Layout {
  uint64_t pos16 = 0;  // next free offset in 16-byte units
  uint32_t alloc(uint64_t bytes) {   // 128-byte aligned, size rounded to 16
    pos16 = (pos16 + 7) & ~(uint64_t)7;
    const uint64_t at = pos16;
    pos16 += (bytes + 15) / 16;
    return (uint32_t)at;
  }
};
int fail(std::string* err, int code, const char* msg) { if (err) *err = msg; return code; }
}  // namespace

int compile_model(const uint8_t* hdr, int hlen, ZModel* M, std::vector<uint8_t>* image, std::string* err) {
  memset(M, 0, sizeof(*M));
  if (!hdr || hlen < 8) return fail(err, ZPQ_E_HEADER, "unexpected end of file");
  const int hsize = hdr[0] + 256 * hdr[1];
  if (hsize + 2 != hlen) return fail(err, ZPQ_E_HEADER, "header size mismatch");
  const int n = hdr[6];
  int pos = 7;
  for (int i = 0; i < n; ++i) {
    if (pos >= hlen) return fail(err, ZPQ_E_HEADER, "unexpected end of file");
    const int t = hdr[pos];
    if (t < 1 || t > 9) return fail(err, ZPQ_E_HEADER, "Invalid component type");
    if (pos + kCompLen[t] > hsize) return fail(err, ZPQ_E_HEADER, "COMP overflows header");
    pos += kCompLen[t];
  }
  if (pos >= hlen || hdr[pos] != 0) return fail(err, ZPQ_E_HEADER, "missing COMP END");
  ++pos;
  const int prog_len = hlen - pos;
  if (prog_len < 1) return fail(err, ZPQ_E_HEADER, "missing HCOMP");
  if (hdr[hlen - 1] != 0) return fail(err, ZPQ_E_HEADER, "missing HCOMP END");
  if (n == 0) return fail(err, ZPQ_E_UNSUPPORTED, "unmodelled (stored) block: not on the context-mixing path");
  const bool wide = n > ZPQ_MAX_LANES;        // more components than lanes: the wide kernel (one thread per block) codes it
  if (hdr[2] > 32) return fail(err, ZPQ_E_COMPONENT, "H too big");
  if (hdr[3] > 32) return fail(err, ZPQ_E_COMPONENT, "M too big");
  if (hdr[2] > 28 || hdr[3] > 31) return fail(err, ZPQ_E_NOMEM, "Out of memory");

  M->n = n; M->hh = hdr[2]; M->hm = hdr[3];
  M->prog_len = prog_len;
  image->assign(ZPQ_GUARD_LO + prog_len + ZPQ_GUARD_HI, 0);
  memcpy(image->data() + ZPQ_GUARD_LO, hdr + pos, prog_len);

  Layout lay;
  uint32_t nf = 0;
  std::vector<ZComp> wcomp(wide ? n : 0);
  std::vector<ZFill> wfills;
  auto fill = [&](uint32_t off16, uint64_t bytes, uint32_t kind, uint32_t value) -> bool {
    ZFill f; f.off16 = off16; f.n16 = (uint32_t)((bytes + 15) / 16); f.kind = kind; f.value = value;
    if (wide) { wfills.push_back(f); ++nf; return true; }
    if (nf >= ZPQ_MAX_FILLS) return false;
    M->fills[nf++] = f;
    return true;
  };
  const uint64_t kMaxTable = (uint64_t)1 << 35;   // 32 GiB per table: beyond any B200 arena

  int ready[256];     // step index after which p[i] is valid (0 = phase A)
  int nsteps = 0;
  const uint8_t* cp = hdr + 7;
  for (int i = 0; i < n; ++i) {
    ZComp& c = wide ? wcomp[i] : M->comp[i];
    memset(&c, 0, sizeof(c));
    c.type = cp[0];
    c.a1 = cp[1];
    c.a2 = kCompLen[cp[0]] > 2 ? cp[2] : 0;
    c.a3 = kCompLen[cp[0]] > 3 ? cp[3] : 0;
    c.a4 = kCompLen[cp[0]] > 4 ? cp[4] : 0;
    c.a5 = kCompLen[cp[0]] > 5 ? cp[5] : 0;
    c.sizebits = cp[1];
    ready[i] = 0;
    bool ok = true;
    switch (cp[0]) {
      case ZT_CONS: break;
      case ZT_CM: {
        if (cp[1] > 32) return fail(err, ZPQ_E_COMPONENT, "max size for CM is 32");
        const uint64_t bytes = (uint64_t)4 << cp[1];
        if (bytes > kMaxTable) return fail(err, ZPQ_E_NOMEM, "Out of memory");
        c.cm_off16 = lay.alloc(bytes);
        c.cm_mask = (uint32_t)(((uint64_t)1 << cp[1]) - 1);
        c.limit = cp[2] * 4;
        ok = fill(c.cm_off16, bytes, ZF_CONST32, 0x80000000u);
        break;
      }
      case ZT_ICM: {
        if (cp[1] > 26) return fail(err, ZPQ_E_COMPONENT, "max size for ICM is 26");
        c.limit = 1023;
        c.cm_off16 = lay.alloc(1024);
        const uint64_t bytes = (uint64_t)64 << cp[1];
        c.ht_off16 = lay.alloc(bytes);
        c.ht_mask = (uint32_t)(bytes - 16);
        c.cm_mask = 255;
        ok = fill(c.cm_off16, 1024, ZF_ICM, 0) && fill(c.ht_off16, bytes, ZF_ZERO, 0);
        break;
      }
      case ZT_MATCH: {
        if (cp[1] > 32 || cp[2] > 32) return fail(err, ZPQ_E_COMPONENT, "max size for MATCH is 32 32");
        const uint64_t ib = (uint64_t)4 << cp[1], bb = (uint64_t)1 << cp[2];
        if (ib > kMaxTable || cp[2] > 31) return fail(err, ZPQ_E_NOMEM, "Out of memory");
        c.cm_off16 = lay.alloc(ib);
        c.ht_off16 = lay.alloc(bb);
        c.cm_mask = (uint32_t)(((uint64_t)1 << cp[1]) - 1);
        c.ht_mask = (uint32_t)(bb - 1);
        ok = fill(c.cm_off16, ib, ZF_ZERO, 0) && fill(c.ht_off16, bb, ZF_BYTE1, 0);
        break;
      }
      case ZT_AVG:
        if (cp[1] >= i) return fail(err, ZPQ_E_COMPONENT, "AVG j >= i");
        if (cp[2] >= i) return fail(err, ZPQ_E_COMPONENT, "AVG k >= i");
        break;
      case ZT_MIX2: {
        if (cp[1] > 32) return fail(err, ZPQ_E_COMPONENT, "max size for MIX2 is 32");
        if (cp[3] >= i) return fail(err, ZPQ_E_COMPONENT, "MIX2 k >= i");
        if (cp[2] >= i) return fail(err, ZPQ_E_COMPONENT, "MIX2 j >= i");
        const uint64_t bytes = (uint64_t)2 << cp[1];
        if (bytes > kMaxTable) return fail(err, ZPQ_E_NOMEM, "Out of memory");
        c.cm_off16 = lay.alloc(bytes);
        c.cm_mask = (uint32_t)(((uint64_t)1 << cp[1]) - 1);
        ok = fill(c.cm_off16, bytes, ZF_CONST32, 0x80008000u);
        break;
      }
      case ZT_MIX: {
        if (cp[1] > 32) return fail(err, ZPQ_E_COMPONENT, "max size for MIX is 32");
        if (cp[2] >= i) return fail(err, ZPQ_E_COMPONENT, "MIX j >= i");
        if (cp[3] < 1 || cp[3] > i - cp[2]) return fail(err, ZPQ_E_COMPONENT, "MIX m not in 1..i-j");
        const uint64_t bytes = ((uint64_t)4 * cp[3]) << cp[1];
        if (cp[1] > 30 || bytes > kMaxTable) return fail(err, ZPQ_E_NOMEM, "Out of memory");
        c.cm_off16 = lay.alloc(bytes);
        c.cm_mask = (uint32_t)(((uint64_t)1 << cp[1]) - 1);
        ok = fill(c.cm_off16, bytes, ZF_CONST32, (uint32_t)(65536 / cp[3]));
        break;
      }
      case ZT_ISSE: {
        if (cp[1] > 32) return fail(err, ZPQ_E_COMPONENT, "max size for ISSE is 32");
        if (cp[2] >= i) return fail(err, ZPQ_E_COMPONENT, "ISSE j >= i");
        if (cp[1] > 26) return fail(err, ZPQ_E_NOMEM, "Out of memory");
        c.cm_off16 = lay.alloc(2048);
        const uint64_t bytes = (uint64_t)64 << cp[1];
        c.ht_off16 = lay.alloc(bytes);
        c.ht_mask = (uint32_t)(bytes - 16);
        c.cm_mask = 511;
        ok = fill(c.cm_off16, 2048, ZF_ISSE, 0) && fill(c.ht_off16, bytes, ZF_ZERO, 0);
        break;
      }
      case ZT_SSE: {
        if (cp[1] > 32) return fail(err, ZPQ_E_COMPONENT, "max size for SSE is 32");
        if (cp[2] >= i) return fail(err, ZPQ_E_COMPONENT, "SSE j >= i");
        if (cp[3] > cp[4] * 4) return fail(err, ZPQ_E_COMPONENT, "SSE start > limit*4");
        if (cp[1] > 27) return fail(err, ZPQ_E_NOMEM, "Out of memory");
        const uint64_t bytes = (uint64_t)128 << cp[1];
        if (bytes > kMaxTable) return fail(err, ZPQ_E_NOMEM, "Out of memory");
        c.cm_off16 = lay.alloc(bytes);
        c.cm_mask = (uint32_t)(((uint64_t)32 << cp[1]) - 1);
        c.limit = cp[4] * 4;
        ok = fill(c.cm_off16, bytes, ZF_SSE, cp[3]);
        break;
      }
      default: return fail(err, ZPQ_E_COMPONENT, "unknown component type");
    }
    if (!ok) return fail(err, ZPQ_E_UNSUPPORTED, "too many tables");

    // ---- dependent part of predict(): place the component in a step ----
    auto new_step = [&](uint8_t kind) -> int {
      if (nsteps >= ZPQ_MAX_STEPS) return -1;
      ZStep& s = M->steps[nsteps];
      s.kind = kind; s.m = 0; s.j0 = 0; s.rate = 0; s.cm_off16 = 0;
      return nsteps++;
    };
    int need = -1;   // -1: not a dependent component
    switch (cp[0]) {
      case ZT_ISSE: need = ready[cp[2]]; break;
      case ZT_SSE: need = ready[cp[2]]; break;
      case ZT_AVG: need = std::max(ready[cp[1]], ready[cp[2]]); break;
      case ZT_MIX2: need = std::max(ready[cp[2]], ready[cp[3]]); break;
      default: break;
    }
    if (wide) { cp += kCompLen[cp[0]]; continue; }      // no step plan: the wide kernel walks the components in order
    if (need >= 0) {
      // first PAR step strictly after `need` (steps are numbered from 1)
      int s = -1;
      for (int k = need; k < nsteps; ++k)
        if (M->steps[k].kind == ZPQ_STEP_PAR) { s = k; break; }
      if (s < 0) s = new_step(ZPQ_STEP_PAR);
      if (s < 0) return fail(err, ZPQ_E_UNSUPPORTED, "too many predict steps");
      ready[i] = s + 1;
      c.step = (uint8_t)(s + 1);
    }
    else if (cp[0] == ZT_MIX) {
      const int s = new_step((uint8_t)i);
      if (s < 0) return fail(err, ZPQ_E_UNSUPPORTED, "too many predict steps");
      M->steps[s].m = cp[3]; M->steps[s].j0 = cp[2]; M->steps[s].rate = cp[4];
      M->steps[s].cm_off16 = c.cm_off16;
      ready[i] = s + 1;
      c.step = (uint8_t)(s + 1);
    }
    cp += kCompLen[cp[0]];
  }
  M->nsteps = nsteps;

  // HCOMP machine memory (ZPAQL::init, libzpaq.cpp:1009-1024)
  const uint64_t mbytes = (uint64_t)1 << M->hm, hbytes = (uint64_t)4 << M->hh;
  M->m_off16 = lay.alloc(mbytes);
  M->h_off16 = lay.alloc(hbytes);
  M->r_off16 = lay.alloc(1024);
  M->m_mask = (uint32_t)(mbytes - 1);
  M->h_mask = (uint32_t)(((uint64_t)1 << M->hh) - 1);
  // M, H and R are adjacent: one zero fill covers the three (padding included)
  {
    const uint64_t from = M->m_off16, to = (uint64_t)M->r_off16 + 64;
    if (!fill((uint32_t)from, (to - from) * 16, ZF_ZERO, 0)) return fail(err, ZPQ_E_UNSUPPORTED, "too many tables");
  }
  if (wide) {
    M->wst_off16 = lay.alloc((uint64_t)n * 32);
    fill(M->wst_off16, (uint64_t)n * 32, ZF_ZERO, 0);
  }
  M->nfills = nf;
  if (wide) {
    if (lay.pos16 >= ((uint64_t)1 << 32)) return fail(err, ZPQ_E_NOMEM, "Out of memory");
    M->arena_bytes = ((lay.pos16 * 16) + 255) & ~(uint64_t)255;
    M->wide = 1;
    while (image->size() & 15) image->push_back(0);
    M->wide_comp_off = (uint32_t)image->size();
    image->insert(image->end(), (const uint8_t*)wcomp.data(), (const uint8_t*)wcomp.data() + wcomp.size() * sizeof(ZComp));
    M->wide_fill_off = (uint32_t)image->size();
    image->insert(image->end(), (const uint8_t*)wfills.data(), (const uint8_t*)wfills.data() + wfills.size() * sizeof(ZFill));
    return ZPQ_OK;
  }
  const uint32_t ring_off16 = lay.alloc((uint64_t)n * 512);   // prediction rings of the team kernel (never needs init)
  if (lay.pos16 >= ((uint64_t)1 << 32)) return fail(err, ZPQ_E_NOMEM, "Out of memory");
  M->arena_bytes = ((lay.pos16 * 16) + 255) & ~(uint64_t)255;

  // ================= fast kernel plan =================
  // ISSE components are grouped by dependency level (one step per level); every other
  // dependent component gets a step of its own.  The plan is only valid (fast.ok) when it
  // fits the register budgets of zpq_kernels2.cu; otherwise the generic kernel codes the block.
  {
    ZFast& F = M->fast;
    memset(&F, 0, sizeof(F));
    bool ok = true;
    int rdy[ZPQ_MAX_LANES];
    int ns = 0;
    const uint8_t* q = hdr + 7;
    for (int i = 0; i < n && ok; ++i) {
      rdy[i] = 0; F.level[i] = 0;
      switch (q[0]) {
        case ZT_ISSE: {
          const int need = rdy[q[2]];
          int at = -1;
          for (int k = need; k < ns; ++k) if ((F.sd[k] & 7) == ZK_ISSE) { at = k; break; }
          if (at < 0) { if (ns >= ZPQ_FAST_STEPS) { ok = false; break; } F.sd[ns] = ZSTEP_PACK(ZK_ISSE, 0, 0, 0, 0, 0); at = ns++; }
          rdy[i] = at + 1; F.level[i] = (uint8_t)(at + 1);
          break;
        }
        case ZT_MIX:
          if (ns >= ZPQ_FAST_STEPS || F.nmix >= ZPQ_FAST_MIX) { ok = false; break; }
          F.mix_off16[F.nmix] = M->comp[i].cm_off16; F.mix_owner[F.nmix] = i; F.mix_m[F.nmix] = q[3];
          F.mix_j0[F.nmix] = q[2]; F.mix_rate[F.nmix] = q[4];
          F.sd[ns++] = ZSTEP_PACK(ZK_MIX, i, q[2], q[3], q[4], F.nmix);
          ++F.nmix; rdy[i] = ns;
          break;
        case ZT_MIX2:
          if (ns >= ZPQ_FAST_STEPS) { ok = false; break; }
          F.sd[ns++] = ZSTEP_PACK(ZK_MIX2, i, q[2], q[3], 0, 0); rdy[i] = ns;
          break;
        case ZT_SSE:
          if (ns >= ZPQ_FAST_STEPS || F.nsse >= ZPQ_FAST_SSE) { ok = false; break; }
          F.sse_off16[F.nsse] = M->comp[i].cm_off16; F.sse_owner[F.nsse] = i; F.sse_mask[F.nsse] = M->comp[i].cm_mask;
          F.sd[ns++] = ZSTEP_PACK(ZK_SSE, i, q[2], 0, 0, F.nsse);
          ++F.nsse; rdy[i] = ns;
          break;
        case ZT_AVG:
          if (ns >= ZPQ_FAST_STEPS) { ok = false; break; }
          F.sd[ns++] = ZSTEP_PACK(ZK_AVG, i, q[1], q[2], q[3], 0); rdy[i] = ns;
          break;
        default: break;
      }
      q += kCompLen[q[0]];
    }
    F.nsteps = ns;
    uint32_t used = 0;
    for (int i = 0; i < n; ++i) {
      const int t = M->comp[i].type;
      F.smem_off[i] = 0;
      if (t == ZT_ICM) { F.smem_off[i] = used; used += 1024; }
      else if (t == ZT_ISSE) { F.smem_off[i] = used; used += 2048; }
    }
    F.pw_bytes = std::max<uint32_t>(16, used);
    if (F.pw_bytes > ZPQ_SMEM_MAX - 38528) ok = false;
    F.ok = ok ? 1 : 0;
  }
  // (3) lane-parallel HCOMP recipes (zpq_hcomp.cpp)
  build_recipes(M, hdr + pos, prog_len);
  if (M->rcp.ok) for (uint32_t l = 0; l < M->n; ++l) if (M->rcp.len[l] > ZPQ_FAST_RCP) M->fast.ok = 0;

  // ================= skewed-pipeline encode plan (depths used by the wavefront encoder) =================
  {
    ZPipe& P = M->pipe;
    memset(&P, 0, sizeof(P));
    bool ok = M->fast.ok && M->rcp.ok && M->hm == 16;   // hm 16: byte offsets below 64 KiB never wrap the M ring
    int dmax = 0, nsp = 0;
    int owner_of[ZPQ_PIPE_STEPS];
    const uint8_t* q = hdr + 7;
    for (int i = 0; i < n && ok; ++i) {
      int d = 0;
      uint32_t word = 0;
      switch (q[0]) {
        case ZT_ISSE: d = P.depth[q[2]] + 1; break;
        case ZT_SSE: d = P.depth[q[2]] + 1; word = ZSTEP_PACK(ZK_SSE, i, q[2], 0, 0, 0); break;
        case ZT_AVG: d = std::max(P.depth[q[1]], P.depth[q[2]]) + 1; word = ZSTEP_PACK(ZK_AVG, i, q[1], q[2], q[3], 0); break;
        case ZT_MIX2: d = std::max(P.depth[q[2]], P.depth[q[3]]) + 1; word = ZSTEP_PACK(ZK_MIX2, i, q[2], q[3], 0, 0); break;
        case ZT_MIX:
          for (int j = 0; j < q[3]; ++j) d = std::max<int>(d, P.depth[q[2] + j] + 1);
          word = ZSTEP_PACK(ZK_MIX, i, q[2], q[3], q[4], 0);
          break;
        default: break;
      }
      P.depth[i] = (uint8_t)d;
      dmax = std::max(dmax, d);
      if (word) {
        if (nsp >= ZPQ_PIPE_STEPS) { ok = false; break; }
        owner_of[nsp] = i;
        P.sd[nsp++] = word;
      }
      q += kCompLen[q[0]];
    }
    if (dmax > 15) ok = false;
    if (ok) {
      // slot numbers of MIX / SSE steps (register slots in the kernel), delays of feeding lanes
      int nm = 0, nss = 0;
      for (int s = 0; s < nsp; ++s) {
        const uint32_t kind = P.sd[s] & 7u;
        const int o = owner_of[s];
        const int j = (int)((P.sd[s] >> 8) & 31u), k = (int)((P.sd[s] >> 13) & 63u);
        if (kind == ZK_MIX) { P.sd[s] |= (uint32_t)nm++ << 27; for (int l = 0; l < k; ++l) P.delay[j + l] |= (uint32_t)(P.depth[o] - P.depth[j + l]) << (4 * s); }
        else if (kind == ZK_SSE) { P.sd[s] |= (uint32_t)nss++ << 27; P.delay[j] |= (uint32_t)(P.depth[o] - P.depth[j]) << (4 * s); }
        else {
          // two inputs; if both are the same lane they share one delay
          P.delay[j] |= (uint32_t)(P.depth[o] - P.depth[j]) << (4 * s);
          if (k != j) P.delay[k] |= (uint32_t)(P.depth[o] - P.depth[k]) << (4 * s);
        }
      }
      // recipe constraints of the skewed evaluation: a hash seeded by another lane must sit exactly one
      // stage behind it, word-model groups likewise; at most one distance source per lane
      const ZRecipes& R = M->rcp;
      for (int l = 0; l < n && ok; ++l) {
        if (R.start[l] == ZB_LANE && P.depth[l] != P.depth[R.base[l]] + 1) ok = false;
        int ndist = 0;
        for (int t = 0; t < R.len[l]; ++t) {
          const uint32_t w = R.ops[t][l];
          if ((w & 15u) == ZR_WORD) { if (!((R.ops[t + 1][l] >> 8) & 1u) && (l == 0 || P.depth[l] != P.depth[l - 1] + 1)) ok = false; ++t; continue; }
          if (((w >> 4) & 15u) == ZS_DIST) { ++ndist; P.has_dist[l] = 1; P.dist_byte[l] = (uint8_t)((w >> 8) & 255u); }
        }
        if (ndist > 1) ok = false;
      }
    }
    P.dmax = dmax; P.nsp = nsp; P.ok = ok ? 1 : 0;
  }

  // ================= three-warp team plan (mixer list used by the wavefront encoder) =================
  {
    ZTeam& T = M->team;
    memset(&T, 0, sizeof(T));
    memset(T.comp_at, 0xff, sizeof(T.comp_at));
    T.ring_off16 = ring_off16;
    bool ok = M->pipe.ok != 0;
    int fill[ZPQ_ROLES] = {0, 0, 0};
    const uint8_t* q = hdr + 7;
    for (int i = 0; i < n && ok; ++i) {
      int role = 2;
      if (q[0] == ZT_ICM || q[0] == ZT_ISSE) role = 0;
      else if (q[0] == ZT_MIX) role = 1;
      T.role_of[i] = (uint8_t)role;
      if (role == 1) { if (T.nmix >= ZPQ_FAST_MIX) ok = false; else T.mix_comp[T.nmix++] = i; }
      else { if (fill[role] >= 32) ok = false; else T.comp_at[role][fill[role]++] = (uint8_t)i; }
      if (q[0] == ZT_SSE) { if (T.nsse >= ZPQ_FAST_SSE) ok = false; else T.sse_comp[T.nsse++] = i; }
      // producer / consumer relations between roles
      auto link = [&](int j) { const int rj = T.role_of[j]; if (rj != role) { T.pmask[role] |= 1u << rj; T.cmask[rj] |= 1u << role; } };
      switch (q[0]) {
        case ZT_ISSE: case ZT_SSE: link(q[2]); break;
        case ZT_AVG: link(q[1]); link(q[2]); break;
        case ZT_MIX2: link(q[2]); link(q[3]); break;
        case ZT_MIX: for (int j = 0; j < q[3]; ++j) link(q[2] + j); break;
        default: break;
      }
      q += kCompLen[q[0]];
    }
    T.coder_role = T.role_of[n - 1];
    // recipe sources must live in the same warp as their consumer; mixer contexts must be pure folds
    for (int l = 0; l < n && ok; ++l) {
      const ZRecipes& R = M->rcp;
      if (R.start[l] == ZB_LANE && T.role_of[R.base[l]] != T.role_of[l]) ok = false;
      if (T.role_of[l] == 1 && R.start[l] != ZB_ZERO) ok = false;
      for (int t = 0; t < R.len[l]; ++t)
        if ((R.ops[t][l] & 15u) == ZR_WORD) {
          if (!((R.ops[t + 1][l] >> 8) & 1u) && (l == 0 || T.role_of[l - 1] != T.role_of[l])) ok = false;
          if (T.role_of[l] == 1) ok = false;
          ++t;
        }
        else if (T.role_of[l] == 1 && ((R.ops[t][l] >> 4) & 15u) == ZS_DIST) ok = false;
    }
    T.ok = ok ? 1 : 0;
  }
  build_wave_plan(M);
  build_hash_plan(M);
  // two-warp decoder: needs the self-contained (flattened) recipes, and every ISSE fed by an ICM or ISSE
  {
    bool ok = M->fast.ok && M->wave.ok;
    uint32_t nlev = 0;
    for (uint32_t s = 0; s < M->fast.nsteps; ++s) if ((M->fast.sd[s] & 7u) == ZK_ISSE) ++nlev;
    for (int i = 0; i < n && ok; ++i)
      if (M->comp[i].type == ZT_ISSE) { const int t = M->comp[M->comp[i].a2].type; if (t != ZT_ICM && t != ZT_ISSE) ok = false; }
    // ISSE levels must be the leading steps (they run in the first warp before anything else is known)
    for (uint32_t s = 0; s < M->fast.nsteps && ok; ++s) if (((M->fast.sd[s] & 7u) == ZK_ISSE) != (s < nlev)) ok = false;
    M->fast.pair_ok = ok ? 1 : 0; M->fast.nlev = nlev;
  }
  // lock-step decoder plan: needs the two-warp preconditions (flattened recipes, ISSE fed by ICM / ISSE), every MIX
  // before every scalar step, and the role counts the kernel is built for
  {
    ZTPlan& T = M->tplan;
    memset(&T, 0, sizeof(T));
    bool ok = M->fast.pair_ok != 0 && M->wave.ok != 0;
    bool seen_scalar = false;
    for (int i = 0; i < n && ok; ++i) {
      const ZComp& c = M->comp[i];
      switch (c.type) {
        case ZT_ICM: case ZT_ISSE:
          if (T.nh >= ZPQ_T_HASH) { ok = false; break; }
          T.hcomp[T.nh++] = (uint8_t)i;
          if (c.type == ZT_ISSE) {
            // the chain warp keeps one running value: an ISSE's input must be an ICM or the ISSE right before it
            if (M->comp[c.a2].type == ZT_ISSE && !(T.nisse > 0 && T.icomp[T.nisse - 1] == c.a2)) ok = false;
            if (T.nisse >= 16) ok = false;
            T.isse_ord[i] = (uint8_t)T.nisse; T.icomp[T.nisse] = (uint8_t)i; T.ijn[T.nisse] = c.a2; ++T.nisse;
          }
          break;
        case ZT_CM: case ZT_MATCH: case ZT_CONS:
          if (T.nmisc >= ZPQ_T_MISC) { ok = false; break; }
          T.mcomp[T.nmisc++] = (uint8_t)i;
          break;
        case ZT_MIX:
          if (seen_scalar || c.a3 > 32) ok = false;        // a mixer fed by a scalar step; more inputs than a component warp has lanes for
          break;
        case ZT_MIX2: case ZT_SSE: case ZT_AVG:
          if (T.nscalar >= ZPQ_T_SCALAR) { ok = false; break; }
          T.scomp[T.nscalar++] = (uint8_t)i;
          seen_scalar = true;
          break;
        default: ok = false;
      }
    }
    // a mixer's inputs must be complete when the mix warp runs: components of warp A kinds, ISSE outputs, earlier mixers
    for (int i = 0; i < n && ok; ++i)
      if (M->comp[i].type == ZT_MIX)
        for (int j = M->comp[i].a2; j < M->comp[i].a2 + M->comp[i].a3; ++j) {
          const uint32_t t = M->comp[j].type;
          if (t == ZT_MIX2 || t == ZT_SSE || t == ZT_AVG) ok = false;
        }
    if (T.nh == 0 || T.nh > 20) ok = false;                 // nh + 4 warps, 768 threads at most
    T.ok = ok ? 1 : 0;
  }
  return ZPQ_OK;
}

}  // namespace zpq
