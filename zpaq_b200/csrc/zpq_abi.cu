// zpq_abi.cu -- C ABI (include/zpq.h) over the kernels: context, batching of
// independent archive blocks, host<->device staging, per-block status.
//
// There is NO CPU implementation of the coding path in this library: every
// entry point that codes data launches zpq_code_kernel and fails with
// ZPQ_E_CUDA when no device is usable.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include <string.h>
#include <map>
#include <string>
#include <thread>
#include <vector>
#include <array>
#include "zpq_host.h"
#include "zpq_kernels.h"

using namespace zpq;

namespace {

struct DevBuf {
  void* p = nullptr; uint64_t cap = 0;
  cudaError_t reserve(uint64_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const uint64_t want = n + (n >> 3) + 4096;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&p, n); if (e == cudaSuccess) cap = n; return e; }
    cap = want;
    return cudaSuccess;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Piece { const uint8_t* p; uint64_t n; };

struct IJob {
  const uint8_t* hdr = nullptr; int hlen = 0;
  std::vector<Piece> in;             // input pieces, concatenated on the device
  std::vector<uint64_t> seg_len;
  uint8_t* out = nullptr; uint64_t out_cap = 0;
  uint64_t out_skip = 0;             // leading device-output bytes not copied to `out`
  std::vector<uint64_t> seg_out;
  uint64_t out_len = 0;
  int status = ZPQ_OK;
  int first_byte = -1;
  // filled by run_batch
  int model = -1;
  uint64_t in_total = 0;
};

}  // namespace

struct zpq_ctx {
  int device = 0, sm_count = 148;
  cudaStream_t stream = nullptr; bool own_stream = true;
  ZTables* d_tables = nullptr;
  void* ws = nullptr; uint64_t ws_bytes = 0; bool own_ws = true;
  DevBuf d_in, d_out, d_meta, d_pbuf, d_cstate, d_prep, d_stage_in, d_stage_out, d_groups[4];
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};      // side streams for concurrent decoder launches of different step signatures
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  std::string err;
  float ms[3] = {0, 0, 0};
  float ms_prep = 0;           // device time of the last analysis pass (SHA-1, gap statistics); added to ms[0] by the whole-block calls
  uint64_t xform_bytes = 0;    // algorithmic HBM bytes of the last pre-processing pass (zpq_bwt_forward)
  float ms_xform = 0;          // device time of the last pre-processing pass (E8E9, BWT: zpq_bwt.cu); added to ms[0] as well
  DevBuf d_xform, d_xws, d_lz; // transformed payloads; workspace of the suffix sorter; LZ77 tables
  uint64_t launches = 0;
  uint32_t tiers[7] = {0, 0, 0, 0, 0, 0, 0};   // blocks of the last batch per kernel tier (zpq_last_tiers)
  int wave_group = 16;         // blocks per component warp of the wavefront encoder (ZPQ_GROUP overrides)
  int wave_chunk = 512;        // bytes per wave of the wavefront encoder (ZPQ_CHUNK overrides; power of two; 512 measured best in round 2: profiles/r02c)
  int wave_variant = 57;       // ZPQ_WAVE_VARIANT: bit 0 byte-ahead MIX rows, bit 1 component threads / MIX warps as two launches, bit 2 mixers as component threads, bit 3 byte-batched MIX warps, bit 4 ICM / ISSE / SSE threads in their own nibble / byte loops, bit 5 compact code (nibble loop rolled, SSE loop unrolled by 2, compact inlined recipe interpreter); 57 measured fastest at the benchmark's size (profiles/r02h_session.log)
  int dec_variant = 31;         // ZPQ_DEC_VARIANT: bit 0 CM in warp A with candidate entries, bit 1 candidate MIX / SSE rows in warp B, bit 2 (only as 7 or 15) slot-scheduled hashes, bit 3 (15) MATCH in warp A, bit 4 (31) register-resident predictions in warp B
  int wave_role_major = 1;     // ZPQ_WAVE_ORDER: 1 = role-major warp order of the wavefront encoder (default), 0 = group-major (round 1)
  bool wave_prof = false;      // ZPQ_WAVE_PROF=1: per-role warp cycles of the wavefront encoder to stderr
  int dec_kernel = 7;           // ZPQ_DEC_KERNEL: 7 = critical-first two-warp decoder (zpq_kernels7.cu, default), 6 = round-1 two-warp decoder, 8 = lock-step decoder with lanes = blocks (zpq_kernels8.cu: correct, measured 2-3x slower, opt-in)
  int kernel_version = 5;      // ZPQ_KERNEL caps the kernel tier: 1 generic, 2 one-warp fast kernel, 5 production (wavefront encoder, two-warp decoder)
};

namespace {

int cuda_fail(zpq_ctx* c, cudaError_t e, const char* what) {
  char buf[256];
  snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
  if (c) c->err = buf;
  cudaGetLastError();
  return ZPQ_E_CUDA;
}
#define CK(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(ctx, e_, what); } while (0)

uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// Run one batch of independent blocks through the GPU.
int run_batch(zpq_ctx* ctx, std::vector<IJob>& jobs, bool decode, bool dev_io) {
  if (!ctx) return ZPQ_E_ARG;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  ctx->ms[0] = ctx->ms[1] = ctx->ms[2] = 0;
  for (int t = 0; t < 7; ++t) ctx->tiers[t] = 0;
  const int nj = (int)jobs.size();
  if (nj == 0) return ZPQ_OK;
  const auto t_begin = std::chrono::steady_clock::now();

  // ---- compile models (shared by blocks with identical headers) ----
  std::vector<ZModel> models;
  std::vector<uint8_t> progs;
  std::map<std::string, int> seen;     // header bytes -> model index, or -1 - (error code) for a header that failed
  for (int i = 0; i < nj; ++i) {
    IJob& J = jobs[i];
    J.status = ZPQ_OK; J.out_len = 0; J.first_byte = -1;
    J.seg_out.assign(J.seg_len.size(), 0);
    J.in_total = 0;
    for (const Piece& pc : J.in) J.in_total += pc.n;
    uint64_t segsum = 0;
    for (uint64_t s : J.seg_len) segsum += s;
    if (!J.hdr || J.hlen <= 0 || J.seg_len.empty() || segsum != J.in_total || (dev_io && J.in.size() != 1)) {
      J.status = ZPQ_E_ARG; continue;
    }
    std::string key((const char*)J.hdr, (size_t)J.hlen);
    auto it = seen.find(key);
    if (it != seen.end()) {
      if (it->second >= 0) J.model = it->second;
      else { J.model = -1; J.status = it->second + 1; }     // undo the -1 - code encoding (codes are negative)
      continue;
    }
    ZModel M; std::vector<uint8_t> image; std::string e;
    const int rc = compile_model(J.hdr, J.hlen, &M, &image, &e);
    if (rc != ZPQ_OK) { J.status = rc; ctx->err = e; seen[key] = rc - 1; J.model = -1; continue; }
    M.prog_off = (uint32_t)progs.size();
    progs.insert(progs.end(), image.begin(), image.end());
    while (progs.size() & 15) progs.push_back(0);
    J.model = (int)models.size();
    seen[key] = J.model;
    models.push_back(M);
  }
  std::vector<int> todo;
  for (int i = 0; i < nj; ++i) if (jobs[i].status == ZPQ_OK && jobs[i].model >= 0) todo.push_back(i);
  if (todo.empty()) return ZPQ_OK;

  // ---- workspace for the model arenas ----
  uint64_t need_all = 0, need_max = 0;
  for (int i : todo) {
    const uint64_t a = models[jobs[i].model].arena_bytes;
    need_all += a; need_max = std::max(need_max, a);
  }
  if (ctx->own_ws && ctx->ws_bytes < need_all) {
    size_t fr = 0, tot = 0;
    CK(cudaMemGetInfo(&fr, &tot), "cudaMemGetInfo");
    const uint64_t avail = (uint64_t)fr + ctx->ws_bytes;
    const uint64_t reserve = ((uint64_t)8 << 30) + ((!decode && ctx->kernel_version >= 5) ? ((uint64_t)14 << 30) : 0);   // staging, CUDA context, prediction streams
    uint64_t want = std::min(need_all, avail > reserve ? avail - reserve : 0);
    if (want > ctx->ws_bytes) {
      if (ctx->ws) { cudaFree(ctx->ws); ctx->ws = nullptr; ctx->ws_bytes = 0; }
      cudaError_t e = cudaMalloc(&ctx->ws, want);
      if (e != cudaSuccess) { cudaGetLastError(); want = std::min(want, std::max(need_max, want / 2)); e = cudaMalloc(&ctx->ws, want); }
      if (e != cudaSuccess) { cuda_fail(ctx, e, "workspace cudaMalloc"); for (int i : todo) jobs[i].status = ZPQ_E_NOMEM; return ZPQ_E_NOMEM; }
      ctx->ws_bytes = want;
    }
  }
  if (ctx->ws_bytes < need_max) {
    for (int i : todo) if (models[jobs[i].model].arena_bytes > ctx->ws_bytes) jobs[i].status = ZPQ_E_NOMEM;
    ctx->err = "Out of memory";
  }

  cudaEvent_t* ev = ctx->ev;
  float ms_init = 0, ms_code = 0, ms_all = 0;
  size_t next = 0;
  while (next < todo.size()) {
    // ---- one wave: as many blocks as the workspace holds ----
    std::vector<int> wave;
    uint64_t used = 0;
    while (next < todo.size()) {
      IJob& J = jobs[todo[next]];
      if (J.status != ZPQ_OK) { ++next; continue; }
      const uint64_t a = models[J.model].arena_bytes;
      if (used + a > ctx->ws_bytes) break;
      used += a; wave.push_back(todo[next]); ++next;
    }
    if (wave.empty()) { if (next < todo.size()) continue; break; }
    const int nb = (int)wave.size();
    // kernel class per block: 0 = wavefront encoder (zpq_kernels5.cu), 3 = step-plan kernels (the two-warp decoders
    // zpq_kernels7.cu / zpq_kernels6.cu, or the one-warp kernel zpq_kernels2.cu for both directions), 4 = generic kernel
    // (zpq_kernels.cu).  ZPQ_KERNEL caps the tier (A/B testing): 1 generic, 2 one-warp fast kernel, 5 production.
    const int kv = ctx->kernel_version;
    auto kclass = [&](int i) -> int {
      const ZModel& Mi = models[jobs[i].model];
      if (Mi.wide) return 5;                                 // more than 32 components: one thread per block (zpq_kernels.cu)
      if (kv == 1) return 4;
      if (!decode && kv >= 5 && Mi.wave.ok) return 0;
      return Mi.fast.ok ? 3 : 4;
    };
    // class first, then model: the wavefront encoder wants 32 blocks of one model per warp
    // (inside the step-plan class: by step signature, so that every signature gets the kernel instantiated for it)
    auto sig_of = [&](int i) -> unsigned long long { return zpq_code2_signature(&models[jobs[i].model].fast); };
    std::stable_sort(wave.begin(), wave.end(), [&](int a, int b) {
      const int ca = kclass(a), cb = kclass(b);
      if (ca != cb) return ca < cb;
      if (ca == 3) { const unsigned long long sa = sig_of(a), sb = sig_of(b); if (sa != sb) return sa < sb; }
      return jobs[a].model < jobs[b].model;
    });
    int ncls[6] = {0, 0, 0, 0, 0, 0};
    for (int i : wave) ++ncls[kclass(i)];

    std::vector<ZBlock> blocks(nb);
    std::vector<uint64_t> seg_len;
    uint64_t in_bytes = 0, out_bytes = 0, arena = 0;
    for (int k = 0; k < nb; ++k) {
      IJob& J = jobs[wave[k]];
      ZBlock& B = blocks[k];
      B.model = (uint32_t)J.model;
      B.seg_first = (uint32_t)seg_len.size(); B.nseg = (uint32_t)J.seg_len.size(); B.pad = 0;
      seg_len.insert(seg_len.end(), J.seg_len.begin(), J.seg_len.end());
      B.arena_off = arena; arena += models[J.model].arena_bytes;
      if (dev_io) {
        B.in_off = (uint64_t)(uintptr_t)J.in[0].p;
        B.out_off = (uint64_t)(uintptr_t)J.out;
        B.out_cap = J.out_cap;
      } else {
        B.in_off = in_bytes; in_bytes = align_up(in_bytes + J.in_total, 16);
        B.out_off = out_bytes; B.out_cap = J.out_cap + J.out_skip;
        out_bytes = align_up(out_bytes + B.out_cap, 16);
      }
    }
    const int nseg = (int)seg_len.size();

    if (!dev_io) {
      // a staging buffer that does not fit fails the blocks of this wave (ZPQ_E_NOMEM), not the batch
      if (ctx->d_in.reserve(in_bytes + 16) != cudaSuccess || ctx->d_out.reserve(out_bytes + 16) != cudaSuccess) {
        cudaGetLastError();
        for (int i : wave) jobs[i].status = ZPQ_E_NOMEM;
        ctx->err = "Out of memory";
        continue;
      }
    }
    // meta: models | blocks | progs | seg_len | seg_out | status | out_len | first | counter
    uint64_t off = 0;
    const uint64_t o_models = off; off = align_up(off + models.size() * sizeof(ZModel), 256);
    const uint64_t o_blocks = off; off = align_up(off + (uint64_t)nb * sizeof(ZBlock), 256);
    const uint64_t o_progs = off;  off = align_up(off + progs.size(), 256);
    const uint64_t o_seglen = off; off = align_up(off + (uint64_t)nseg * 8, 256);
    const uint64_t o_segout = off; off = align_up(off + (uint64_t)nseg * 8, 256);
    const uint64_t o_status = off; off = align_up(off + (uint64_t)nb * 4, 256);
    const uint64_t o_outlen = off; off = align_up(off + (uint64_t)nb * 8, 256);
    const uint64_t o_counter = off; off = align_up(off + 4, 256);
    CK(ctx->d_meta.reserve(off), "meta cudaMalloc");
    uint8_t* dm = (uint8_t*)ctx->d_meta.p;
    cudaStream_t st = ctx->stream;

    CK(cudaMemcpyAsync(dm + o_models, models.data(), models.size() * sizeof(ZModel), cudaMemcpyHostToDevice, st), "H2D models");
    CK(cudaMemcpyAsync(dm + o_blocks, blocks.data(), (size_t)nb * sizeof(ZBlock), cudaMemcpyHostToDevice, st), "H2D blocks");
    CK(cudaMemcpyAsync(dm + o_progs, progs.data(), progs.size(), cudaMemcpyHostToDevice, st), "H2D programs");
    CK(cudaMemcpyAsync(dm + o_seglen, seg_len.data(), (size_t)nseg * 8, cudaMemcpyHostToDevice, st), "H2D segments");
    CK(cudaMemsetAsync(dm + o_segout, 0, (size_t)(o_counter + 256 - o_segout), st), "memset");
    if (!dev_io) {
      for (int k = 0; k < nb; ++k) {
        IJob& J = jobs[wave[k]];
        uint64_t at = blocks[k].in_off;
        for (const Piece& pc : J.in) {
          if (pc.n) CK(cudaMemcpyAsync((uint8_t*)ctx->d_in.p + at, pc.p, pc.n, cudaMemcpyDefault, st), "input staging");      // a piece may be host or device memory
          at += pc.n;
        }
      }
    }

    ZKernelArgs A;
    A.models = (const ZModel*)(dm + o_models);
    A.blocks = (const ZBlock*)(dm + o_blocks);
    A.progs = dm + o_progs;
    A.seg_len = (const uint64_t*)(dm + o_seglen);
    A.seg_out = (uint64_t*)(dm + o_segout);
    A.in = dev_io ? (const uint8_t*)nullptr : (const uint8_t*)ctx->d_in.p;
    A.out = dev_io ? (uint8_t*)nullptr : (uint8_t*)ctx->d_out.p;
    A.arena = (uint8_t*)ctx->ws;
    A.status = (int32_t*)(dm + o_status);
    A.out_len = (uint64_t*)(dm + o_outlen);
    A.counter = (unsigned int*)(dm + o_counter);
    A.tables = ctx->d_tables;
    A.nblk = nb;

    CK(cudaEventRecord(ev[1], st), "event");
    CK(zpq_launch_init(&A, nb, st), "zpq_init_kernel launch");
    CK(cudaEventRecord(ev[2], st), "event");
    {
      int first = 0;
      for (int cls = 0; cls < 6; ++cls) {
        const int cnt = ncls[cls];
        if (cnt == 0) continue;
        ZKernelArgs Ak = A;
        Ak.blocks = A.blocks + first; Ak.status = A.status + first; Ak.out_len = A.out_len + first;
        Ak.counter = A.counter + cls; Ak.nblk = cnt;
        if (cls == 5) { CK(zpq_launch_wide(&Ak, cnt, decode ? 1 : 0, st), "zpq_wide_kernel launch"); ++ctx->launches; ctx->tiers[5] += (uint32_t)cnt; }
        else if (cls == 4) { CK(zpq_launch_code(&Ak, cnt, decode ? 1 : 0, ctx->sm_count, st), "zpq_code_kernel launch"); ++ctx->launches; ctx->tiers[4] += (uint32_t)cnt; }
        else if (cls == 0) {
          // ---- wavefront encoder: groups of <= 32 blocks of one model, one launch per wave ----
          const int GS = ctx->wave_group;
          ctx->tiers[0] += (uint32_t)cnt;
          std::vector<uint32_t> gfirst, gcount, bgroup(cnt), blane(cnt);
          int ncs = 2; uint64_t maxN = 0; uint32_t dmaxc = 0;
          for (int k = 0; k < cnt; ++k) {
            const IJob& J = jobs[wave[first + k]];
            if (k == 0 || J.model != jobs[wave[first + k - 1]].model || (int)gcount.back() == GS) { gfirst.push_back((uint32_t)k); gcount.push_back(0); }
            bgroup[k] = (uint32_t)gfirst.size() - 1; blane[k] = gcount.back()++;
            ncs = std::max(ncs, (int)models[J.model].n + 1);
            maxN = std::max(maxN, J.in_total);
            dmaxc = std::max(dmaxc, models[J.model].wave.dcoder);
          }
          const int ngroups = (int)gfirst.size();
          int CB = ctx->wave_chunk;
          size_t fr = 0, tot = 0;
          CK(cudaMemGetInfo(&fr, &tot), "cudaMemGetInfo");
          while (CB > 128 && (uint64_t)ngroups * ncs * 16 * CB * 16 * GS > (uint64_t)fr + ctx->d_pbuf.cap - ((uint64_t)1 << 30)) CB >>= 1;
          const uint64_t pbytes = (uint64_t)ngroups * ncs * 16 * CB * 16 * GS, cbytes = (uint64_t)cnt * ncs * 64;
          CK(ctx->d_pbuf.reserve(pbytes), "prediction stream cudaMalloc");
          CK(ctx->d_cstate.reserve(cbytes + (uint64_t)(ngroups * 2 + cnt * 2) * 4 + 2048), "wave state cudaMalloc");
          uint8_t* dc = (uint8_t*)ctx->d_cstate.p;
          uint32_t* d_gfirst = (uint32_t*)(dc + align_up(cbytes, 256));
          uint32_t* d_gcount = d_gfirst + ngroups;
          uint32_t* d_bgroup = d_gcount + ngroups;
          uint32_t* d_blane = d_bgroup + cnt;
          unsigned long long* d_prof = (unsigned long long*)(((uintptr_t)(d_blane + cnt) + 15) & ~(uintptr_t)15);
          CK(cudaMemsetAsync(dc, 0, cbytes, st), "memset wave state");
          CK(cudaMemcpyAsync(d_gfirst, gfirst.data(), (size_t)ngroups * 4, cudaMemcpyHostToDevice, st), "H2D groups");
          CK(cudaMemcpyAsync(d_gcount, gcount.data(), (size_t)ngroups * 4, cudaMemcpyHostToDevice, st), "H2D groups");
          CK(cudaMemcpyAsync(d_bgroup, bgroup.data(), (size_t)cnt * 4, cudaMemcpyHostToDevice, st), "H2D groups");
          CK(cudaMemcpyAsync(d_blane, blane.data(), (size_t)cnt * 4, cudaMemcpyHostToDevice, st), "H2D groups");
          CK(cudaStreamSynchronize(st), "sync");     // host vectors above go out of scope
          ZWaveArgs Wv;
          Wv.A = Ak; Wv.pbuf = (uint4*)ctx->d_pbuf.p; Wv.cstate = (uint32_t*)dc;
          Wv.group_first = d_gfirst; Wv.group_count = d_gcount; Wv.blk_group = d_bgroup; Wv.blk_lane = d_blane;
          Wv.chunk_bytes = CB; Wv.ncs = ncs; Wv.ngroups = ngroups; Wv.gsize = GS;
          Wv.prof = nullptr;
          // ZPQ_WAVE_VARIANT bit 2: mixers as component threads, when every MIX of every model in the launch qualifies
          Wv.hash_loop = (ctx->wave_variant & 16) ? 1 : 0;       // bit 4: nibble-unrolled ICM / ISSE loop
          Wv.mix_batch = (ctx->wave_variant & 8) ? 1 : 0;        // bit 3: byte-batched MIX warps
          Wv.mix_threads = (ctx->wave_variant & 4) ? 1 : 0;
          Wv.role_major = ctx->wave_role_major;
          for (int k = first; k < first + cnt && Wv.mix_threads; ++k) {
            const ZModel& Mk = models[jobs[wave[k]].model];
            for (uint32_t ci = 0; ci < Mk.n; ++ci)
              if (Mk.comp[ci].type == ZT_MIX && !((Mk.comp[ci].a5 & 255u) == 255u && Mk.comp[ci].cm_mask >= 255u)) Wv.mix_threads = 0;
          }
          if (ctx->wave_prof) { CK(cudaMemsetAsync(d_prof, 0, 64 * 8, st), "memset prof"); Wv.prof = d_prof; }
          const int nwaves = (int)((maxN + CB - 1) / CB) + (int)dmaxc + 1;
          for (int w = 0; w < std::max(nwaves, (int)dmaxc + 1); ++w) {
            Wv.wave = w;
            CK(zpq_launch_wave(&Wv, st, ctx->wave_variant), "zpq_wave_kernel launch");
            ctx->launches += (ctx->wave_variant & 2) ? 2 : 1;
          }
          if (ctx->wave_prof) {
            unsigned long long hp[64];
            CK(cudaMemcpyAsync(hp, d_prof, sizeof hp, cudaMemcpyDeviceToHost, st), "D2H prof");
            CK(cudaStreamSynchronize(st), "sync");
            static const char* const rn[14] = {"none", "const", "cm", "icm", "match", "avg", "mix2", "mix(thread)", "isse", "sse", "coder", "mix warp 0", "mix warp 1", "mix warp 2"};
            for (int r = 0; r < 14; ++r)
              if (hp[2 * r + 1]) fprintf(stderr, "[zpq wave prof] %-12s warps %9llu  cycles/warp %10.0f  cycles/bit %8.1f\n", rn[r], hp[2 * r + 1],
                                         (double)hp[2 * r] / (double)hp[2 * r + 1], (double)hp[2 * r] / (double)hp[2 * r + 1] / (8.0 * CB));
          }
        }
        else {
          // Runs of one step signature are launched side by side on auxiliary streams (a launch is bound by the latency
          // of its longest block, not by the number of blocks: one after the other would multiply that time).
          std::vector<std::pair<int, int>> runs;
          for (int k = first; k < first + cnt;) {
            int e = k + 1;
            while (e < first + cnt && sig_of(wave[e]) == sig_of(wave[k])) ++e;
            runs.push_back(std::make_pair(k, e - k));
            k = e;
          }
          const int first_all = first, cnt_all = cnt;
          for (size_t rix = 0; rix < runs.size(); ++rix) {
          const int first = runs[rix].first, cnt = runs[rix].second;
          ZKernelArgs Ak = A;
          Ak.blocks = A.blocks + first; Ak.status = A.status + first; Ak.out_len = A.out_len + first;
          Ak.counter = A.counter + 8 + (rix % 48); Ak.nblk = cnt;
          cudaStream_t st = ctx->stream;
          if (rix > 0 && runs.size() > 1) {
            const int ax = (int)((rix - 1) % 3);
            if (!ctx->aux[ax]) CK(cudaStreamCreateWithFlags(&ctx->aux[ax], cudaStreamNonBlocking), "aux stream");
            if (!ctx->ev_fork) { CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming), "event"); }
            CK(cudaEventRecord(ctx->ev_fork, ctx->stream), "event");
            st = ctx->aux[ax];
            CK(cudaStreamWaitEvent(st, ctx->ev_fork, 0), "fork");
          }
          // one CTA per SM; W blocks in flight share the 227 KB carve-out
          uint32_t pw = 16;
          for (int k = first; k < first + cnt; ++k) pw = std::max(pw, models[jobs[wave[k]].model].fast.pw_bytes);
          const uint64_t room = ZPQ_SMEM_MAX - (uint64_t)zpq_code2_static_smem();
          {
            const int warps = (int)std::max<uint64_t>(1, std::min<uint64_t>(8, room / pw));
            {
              unsigned long long sig = zpq_code2_signature(&models[jobs[wave[first]].model].fast);
              for (int k = first + 1; k < first + cnt; ++k)
                if (zpq_code2_signature(&models[jobs[wave[k]].model].fast) != sig) { sig = 0; break; }
              bool pair = decode && kv >= 5;
              for (int k = first; k < first + cnt && pair; ++k) pair = models[jobs[wave[k]].model].fast.pair_ok != 0;
              bool hp_all = true;
              for (int k = first; k < first + cnt; ++k) if (!models[jobs[wave[k]].model].hplan.ok) { hp_all = false; break; }
              bool tp_all = decode && kv >= 5 && ctx->dec_kernel >= 8;
              uint32_t nh_max = 0;
              for (int k = first; k < first + cnt && tp_all; ++k) {
                const ZModel& Mk = models[jobs[wave[k]].model];
                tp_all = Mk.tplan.ok != 0; nh_max = std::max(nh_max, Mk.tplan.nh);
              }
              if (tp_all) {
                // lanes = blocks: one CTA decodes up to eight blocks of one model in lock step (zpq_kernels8.cu)
                const uint64_t room8 = ZPQ_SMEM_MAX - (uint64_t)zpq_dec8_static_smem() - (uint64_t)zpq_dec8_xchg_bytes();
                const int nbk = (int)std::max<uint64_t>(1, std::min<uint64_t>(8, room8 / pw));
                std::vector<uint32_t> gf, gc;
                for (int k = first; k < first + cnt; ++k) {
                  if (gf.empty() || jobs[wave[k]].model != jobs[wave[k - 1]].model || (int)gc.back() == nbk) { gf.push_back((uint32_t)(k - first)); gc.push_back(0); }
                  ++gc.back();
                }
                const int ng = (int)gf.size();
                DevBuf& gb = ctx->d_groups[rix % 4];
                CK(gb.reserve((uint64_t)ng * 8 + 64), "group table cudaMalloc");
                CK(cudaMemcpyAsync(gb.p, gf.data(), (size_t)ng * 4, cudaMemcpyHostToDevice, st), "H2D groups");
                CK(cudaMemcpyAsync((uint8_t*)gb.p + (size_t)ng * 4, gc.data(), (size_t)ng * 4, cudaMemcpyHostToDevice, st), "H2D groups");
                CK(cudaStreamSynchronize(st), "groups");          // the host vectors go out of scope
                CK(zpq_launch_dec8(&Ak, (const uint32_t*)gb.p, (const uint32_t*)((uint8_t*)gb.p + (size_t)ng * 4), ng, nbk, (int)pw,
                                   (int)nh_max + 4, ctx->sm_count, st), "zpq_dec_kernel8 launch");
                ctx->tiers[6] += (uint32_t)cnt;
              }
              else if (pair && hp_all && ctx->dec_kernel >= 7 && zpq_dec7_takes(sig)) {
                // two warps per block, critical-first ordering (zpq_kernels7.cu)
                const uint64_t room7 = ZPQ_SMEM_MAX - (uint64_t)zpq_dec7_static_smem();
                const int teams = (int)std::max<uint64_t>(1, std::min<uint64_t>(7, room7 / (pw + (uint64_t)zpq_dec7_ctrl_bytes())));
                CK(zpq_launch_dec7(&Ak, cnt, ctx->sm_count, teams, (int)pw, sig, st), "zpq_dec_kernel7 launch");
                ctx->tiers[1] += (uint32_t)cnt;
              }
              else if (pair) {
                // two warps per block (zpq_kernels6.cu)
                const uint64_t room6 = ZPQ_SMEM_MAX - (uint64_t)zpq_dec6_static_smem();
                const int teams = (int)std::max<uint64_t>(1, std::min<uint64_t>(7, room6 / (pw + (uint64_t)zpq_dec6_ctrl_bytes())));
                int dv = ctx->dec_variant;
                if (dv & 4) for (int k = first; k < first + cnt; ++k) if (!models[jobs[wave[k]].model].hplan.ok) { dv &= 3; break; }   // 7 / 15 need the hash plan
                CK(zpq_launch_dec6(&Ak, cnt, ctx->sm_count, teams, (int)pw, sig, dv, st), "zpq_dec_kernel6 launch");
                ctx->tiers[2] += (uint32_t)cnt;
              }
              else { CK(zpq_launch_code2(&Ak, cnt, decode ? 1 : 0, ctx->sm_count, warps, (int)pw, sig, st), "zpq_code_kernel2 launch"); ctx->tiers[3] += (uint32_t)cnt; }
            }
          }
          ++ctx->launches;
          if (st != ctx->stream) {
            if (!ctx->ev_join[rix % 3]) CK(cudaEventCreateWithFlags(&ctx->ev_join[rix % 3], cudaEventDisableTiming), "event");
            CK(cudaEventRecord(ctx->ev_join[rix % 3], st), "event");
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[rix % 3], 0), "join");
          }
          }
          (void)first_all; (void)cnt_all;
        }
        first += cnt;
      }
    }
    CK(cudaEventRecord(ev[3], st), "event");
    ++ctx->launches;   // init kernel

    std::vector<int32_t> status(nb);
    std::vector<uint64_t> out_len(nb), seg_out(nseg);
    CK(cudaMemcpyAsync(status.data(), dm + o_status, (size_t)nb * 4, cudaMemcpyDeviceToHost, st), "D2H status");
    CK(cudaMemcpyAsync(out_len.data(), dm + o_outlen, (size_t)nb * 8, cudaMemcpyDeviceToHost, st), "D2H lengths");
    CK(cudaMemcpyAsync(seg_out.data(), dm + o_segout, (size_t)nseg * 8, cudaMemcpyDeviceToHost, st), "D2H segments");
    CK(cudaStreamSynchronize(st), "zpq_code_kernel");
    for (int k = 0; k < nb; ++k) {
      IJob& J = jobs[wave[k]];
      J.status = status[k];
      J.out_len = out_len[k];
      for (uint32_t s = 0; s < blocks[k].nseg; ++s) J.seg_out[s] = seg_out[blocks[k].seg_first + s];
    }
    if (!dev_io) {
      // first decoded byte of each block (the post-processor flag) + payload copies
      for (int k = 0; k < nb; ++k) {
        IJob& J = jobs[wave[k]];
        if (J.status != ZPQ_OK && J.status != ZPQ_E_OUTPUT_FULL) continue;
        const uint64_t have = std::min(J.out_len, blocks[k].out_cap);
        const uint8_t* src = (const uint8_t*)ctx->d_out.p + blocks[k].out_off;
        if (J.out_skip && have) {
          uint8_t fb = 0;
          CK(cudaMemcpyAsync(&fb, src, 1, cudaMemcpyDeviceToHost, st), "D2H flag");
          CK(cudaStreamSynchronize(st), "sync");
          J.first_byte = fb;
        }
        if (have > J.out_skip && J.out)
          CK(cudaMemcpyAsync(J.out, src + J.out_skip, have - J.out_skip, cudaMemcpyDefault, st), "output copy");          // host or device destination
      }
    }
    CK(cudaStreamSynchronize(st), "stream sync");
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, ev[1], ev[2]);
    cudaEventElapsedTime(&b, ev[2], ev[3]);
    ms_init += a; ms_code += b;
  }
  ms_all = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  ctx->ms[0] = ms_init; ctx->ms[1] = ms_code; ctx->ms[2] = ms_all;
  return ZPQ_OK;
}

int jobs_from_abi(int nblk, zpq_job* jobs, std::vector<IJob>* out) {
  if (nblk < 0 || (nblk > 0 && !jobs)) return ZPQ_E_ARG;
  out->resize(nblk);
  for (int i = 0; i < nblk; ++i) {
    IJob& J = (*out)[i];
    J.hdr = jobs[i].hdr; J.hlen = jobs[i].hlen;
    J.in.push_back(Piece{jobs[i].data, jobs[i].len});
    if (jobs[i].seg_len && jobs[i].nseg > 0) J.seg_len.assign(jobs[i].seg_len, jobs[i].seg_len + jobs[i].nseg);
    else J.seg_len.assign(1, jobs[i].len);
    J.out = jobs[i].out; J.out_cap = jobs[i].out_cap;
  }
  return ZPQ_OK;
}

void jobs_to_abi(const std::vector<IJob>& in, zpq_job* jobs) {
  for (size_t i = 0; i < in.size(); ++i) {
    jobs[i].status = in[i].status;
    jobs[i].out_len = in[i].out_len;
    if (jobs[i].seg_out) for (size_t s = 0; s < in[i].seg_out.size(); ++s) jobs[i].seg_out[s] = in[i].seg_out[s];
  }
}

int abi_batch(zpq_ctx* ctx, int nblk, zpq_job* jobs, bool decode, bool dev_io) {
  if (!ctx) return ZPQ_E_ARG;
  std::vector<IJob> ij;
  int rc = jobs_from_abi(nblk, jobs, &ij);
  if (rc != ZPQ_OK) return rc;
  rc = run_batch(ctx, ij, decode, dev_io);
  jobs_to_abi(ij, jobs);
  return rc;
}

}  // namespace

extern "C" {

int zpq_abi_version(void) { return ZPQ_ABI_VERSION; }

const char* zpq_strerror(int s) {
  switch (s) {
    case ZPQ_OK: return "ok";
    case ZPQ_E_HEADER: return "invalid block header";
    case ZPQ_E_COMPONENT: return "invalid component";
    case ZPQ_E_VM: return "ZPAQL execution error";
    case ZPQ_E_OUTPUT_FULL: return "output buffer too small";
    case ZPQ_E_CORRUPT: return "archive corrupted";
    case ZPQ_E_EOF: return "unexpected end of file";
    case ZPQ_E_NOMEM: return "Out of memory";
    case ZPQ_E_UNSUPPORTED: return "not supported by the B200 context-mixing core";
    case ZPQ_E_CUDA: return "CUDA error";
    case ZPQ_E_ARG: return "bad argument";
    case ZPQ_E_EOS: return "decoding end of stream";
    case ZPQ_E_CHECKSUM: return "checksum mismatch";
    case ZPQ_E_METHOD: return "Unsupported method";
    default: return "unknown status";
  }
}

int zpq_create(int device, zpq_ctx** out) {
  if (!out) return ZPQ_E_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) { cudaGetLastError(); return ZPQ_E_CUDA; }
  zpq_ctx* ctx = new zpq_ctx();
  ctx->device = device;
  if (const char* gs = getenv("ZPQ_GROUP")) { const int v = atoi(gs); if (v >= 1 && v <= 32) ctx->wave_group = v; }
  if (const char* cb = getenv("ZPQ_CHUNK")) { const int v = atoi(cb); if (v >= 128 && v <= 65536 && (v & (v - 1)) == 0) ctx->wave_chunk = v; }
  if (const char* wv = getenv("ZPQ_WAVE_VARIANT")) { const int v = atoi(wv); if (v >= 0 && v <= 63) ctx->wave_variant = v; }
  if (const char* dv = getenv("ZPQ_DEC_VARIANT")) { const int v = atoi(dv); if ((v >= 0 && v <= 3) || v == 7 || v == 15 || v == 31) ctx->dec_variant = v; }
  if (const char* dk = getenv("ZPQ_DEC_KERNEL")) { const int v = atoi(dk); if (v >= 6 && v <= 8) ctx->dec_kernel = v; }
  if (const char* wo = getenv("ZPQ_WAVE_ORDER")) ctx->wave_role_major = atoi(wo) != 0;
  if (const char* wp = getenv("ZPQ_WAVE_PROF")) ctx->wave_prof = atoi(wp) != 0;
  if (const char* kv = getenv("ZPQ_KERNEL")) { const int v = atoi(kv); if (v >= 1 && v <= 5) ctx->kernel_version = v; }
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return ZPQ_E_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return ZPQ_E_CUDA; }
  for (int i = 0; i < 4; ++i) if (cudaEventCreate(&ctx->ev[i]) != cudaSuccess) { delete ctx; return ZPQ_E_CUDA; }
  ZTables* T = new ZTables();
  build_tables(T);
  e = cudaMalloc((void**)&ctx->d_tables, sizeof(ZTables));
  if (e == cudaSuccess) e = cudaMemcpy(ctx->d_tables, T, sizeof(ZTables), cudaMemcpyHostToDevice);
  delete T;
  if (e != cudaSuccess) { cudaGetLastError(); zpq_destroy(ctx); return ZPQ_E_CUDA; }
  *out = ctx;
  return ZPQ_OK;
}

void zpq_destroy(zpq_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  ctx->d_in.release(); ctx->d_out.release(); ctx->d_meta.release(); ctx->d_pbuf.release(); ctx->d_cstate.release(); ctx->d_prep.release(); ctx->d_xform.release(); ctx->d_xws.release(); ctx->d_lz.release(); ctx->d_stage_in.release(); ctx->d_stage_out.release(); for (int i = 0; i < 4; ++i) ctx->d_groups[i].release();
  if (ctx->own_ws && ctx->ws) cudaFree(ctx->ws);
  if (ctx->d_tables) cudaFree(ctx->d_tables);
  for (int i = 0; i < 4; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 3; ++i) { if (ctx->aux[i]) { cudaStreamSynchronize(ctx->aux[i]); cudaStreamDestroy(ctx->aux[i]); } if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]); }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* zpq_last_error(const zpq_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }

int zpq_set_workspace(zpq_ctx* ctx, void* p, uint64_t bytes) {
  if (!ctx) return ZPQ_E_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->own_ws && ctx->ws) cudaFree(ctx->ws);
  ctx->ws = p; ctx->ws_bytes = p ? bytes : 0; ctx->own_ws = (p == nullptr);
  if (!p) { ctx->ws = nullptr; ctx->ws_bytes = 0; }
  return ZPQ_OK;
}

int zpq_set_stream(zpq_ctx* ctx, void* stream) {
  if (!ctx) return ZPQ_E_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->own_stream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); ctx->stream = nullptr; }
  if (stream) { ctx->stream = (cudaStream_t)stream; ctx->own_stream = false; }
  else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return ZPQ_E_CUDA;
    ctx->own_stream = true;
  }
  return ZPQ_OK;
}

int zpq_model_bytes(const uint8_t* hdr, int hlen, uint64_t* bytes) {
  ZModel M; std::vector<uint8_t> img; std::string e;
  const int rc = compile_model(hdr, hlen, &M, &img, &e);
  if (rc == ZPQ_OK && bytes) *bytes = M.arena_bytes;
  return rc;
}

// Diagnostics: info[0]=n, [1]=v2 steps, [2]=per-warp shared bytes, [3]=recipes ok, [4]=recipe rounds,
// [5]=fast kernel usable, [6]=v1 steps, [7]=HCOMP bytes
int zpq_model_describe(const uint8_t* hdr, int hlen, uint32_t info[8]) {
  ZModel* M = new ZModel(); std::vector<uint8_t> img; std::string e;
  const int rc = compile_model(hdr, hlen, M, &img, &e);
  if (rc == ZPQ_OK && info) {
    info[0] = M->n; info[1] = M->fast.nsteps; info[2] = M->fast.pw_bytes; info[3] = M->rcp.ok; info[4] = M->rcp.nrounds;
    info[5] = M->fast.ok; info[6] = M->nsteps; info[7] = M->prog_len;
  }
  delete M;
  return rc;
}

int zpq_model_from_method(const char* method, const uint8_t* data, uint64_t n, uint8_t* hdr, int cap, int* hlen) {
  if (!method || !hdr || !hlen) return ZPQ_E_ARG;
  std::vector<uint8_t> h; std::string e;
  const int rc = header_from_method(method, data, n, &h, &e);
  if (rc != ZPQ_OK) return rc;
  if ((int)h.size() > cap) return ZPQ_E_OUTPUT_FULL;
  memcpy(hdr, h.data(), h.size());
  *hlen = (int)h.size();
  return ZPQ_OK;
}

int zpq_header_from_stats(const char* method, uint64_t n, const int32_t gaps[2], uint8_t* hdr, int cap, int* hlen) {
  if (!method || !hdr || !hlen) return ZPQ_E_ARG;
  std::vector<uint8_t> h; std::string e;
  const int g[2] = {gaps ? gaps[0] : 0, gaps ? gaps[1] : 0};
  const int rc = header_from_stats(method, n, g, &h, &e);
  if (rc != ZPQ_OK) return rc;
  if ((int)h.size() > cap) return ZPQ_E_OUTPUT_FULL;
  memcpy(hdr, h.data(), h.size());
  *hlen = (int)h.size();
  return ZPQ_OK;
}

// header + pre-processing kind + model prefix of a method (what compress_core builds per block)
static int plan_block(const char* method, uint64_t n, const int g[2], std::vector<uint8_t>* hdr, int* prep, std::vector<uint8_t>* prefix, std::string* err,
                      int* size_class_out = nullptr, int* lz_out = nullptr) {
  int size_class = 0;
  int lz[5] = {0, 0, 0, 0, 0};
  const int rc = block_plan(method, n, g, hdr, prep, &size_class, lz, err);
  if (size_class_out) *size_class_out = size_class;
  if (lz_out) for (int i = 0; i < 5; ++i) lz_out[i] = lz[i];
  if (rc != ZPQ_OK) return rc;
  prefix->assign(1, 0);                                  // PASS
  if (*prep != 0) {
    std::vector<uint8_t> code;
    const int rp = pcomp_program(*prep, size_class, lz[0], &code);
    if (rp != ZPQ_OK) { *err = "no post-processor program for this pre-processing"; return rp; }
    pcomp_prefix(code, prefix);
  }
  return ZPQ_OK;
}

int zpq_block_plan(const char* method, uint64_t n, const int32_t gaps[2], uint8_t* hdr, int hdr_cap, int* hlen,
                   int* prep, uint8_t* prefix, int prefix_cap, int* prefix_len) {
  if (!method || !hdr || !hlen || !prep || !prefix || !prefix_len) return ZPQ_E_ARG;
  std::vector<uint8_t> h, pre; std::string e;
  const int g[2] = {gaps ? gaps[0] : 0, gaps ? gaps[1] : 0};
  const int rc = plan_block(method, n, g, &h, prep, &pre, &e);
  if (rc != ZPQ_OK) return rc;
  if ((int)h.size() > hdr_cap || (int)pre.size() > prefix_cap) return ZPQ_E_OUTPUT_FULL;
  memcpy(hdr, h.data(), h.size()); *hlen = (int)h.size();
  memcpy(prefix, pre.data(), pre.size()); *prefix_len = (int)pre.size();
  return ZPQ_OK;
}

int zpq_encode_blocks(zpq_ctx* ctx, int nblk, zpq_job* jobs) { return abi_batch(ctx, nblk, jobs, false, false); }
int zpq_decode_blocks(zpq_ctx* ctx, int nblk, zpq_job* jobs) { return abi_batch(ctx, nblk, jobs, true, false); }
int zpq_encode_blocks_dev(zpq_ctx* ctx, int nblk, zpq_job* jobs) { return abi_batch(ctx, nblk, jobs, false, true); }
int zpq_decode_blocks_dev(zpq_ctx* ctx, int nblk, zpq_job* jobs) { return abi_batch(ctx, nblk, jobs, true, true); }

int zpq_last_timing(const zpq_ctx* ctx, float ms[3]) {
  if (!ctx || !ms) return ZPQ_E_ARG;
  ms[0] = ctx->ms[0]; ms[1] = ctx->ms[1]; ms[2] = ctx->ms[2];
  return ZPQ_OK;
}

int zpq_last_prep_timing(const zpq_ctx* ctx, float ms[2]) {
  if (!ctx || !ms) return ZPQ_E_ARG;
  ms[0] = ctx->ms_prep; ms[1] = ctx->ms_xform;
  return ZPQ_OK;
}

uint64_t zpq_last_transform_bytes(const zpq_ctx* ctx) { return ctx ? ctx->xform_bytes : 0; }

uint64_t zpq_launch_count(const zpq_ctx* ctx) { return ctx ? ctx->launches : 0; }

int zpq_last_tiers(const zpq_ctx* ctx, uint32_t tiers[7]) {
  if (!ctx || !tiers) return ZPQ_E_ARG;
  for (int t = 0; t < 7; ++t) tiers[t] = ctx->tiers[t];
  return ZPQ_OK;
}

int zpq_postprocess_block(const uint8_t* hdr, int hlen, const uint8_t* model_out, const uint64_t* seg_len, int nseg,
                          uint8_t* out, uint64_t out_cap, uint64_t* out_len, uint64_t* seg_out) {
  if (!hdr || hlen < 7 || !model_out || !seg_len || nseg < 1 || !out_len) return ZPQ_E_ARG;
  std::vector<uint8_t> payload; std::vector<uint64_t> segs; std::string e;
  const int rc = postprocess_block(hdr[4], hdr[5], model_out, seg_len, nseg, &payload, &segs, &e);
  if (rc != ZPQ_OK) return rc;
  *out_len = payload.size();
  if (seg_out) for (int i = 0; i < nseg; ++i) seg_out[i] = segs[i];
  if (payload.size() > out_cap || (!out && !payload.empty())) return ZPQ_E_OUTPUT_FULL;
  if (!payload.empty()) memcpy(out, payload.data(), payload.size());
  return ZPQ_OK;
}

uint64_t zpq_compress_bound(uint64_t len) { return len + len / 8 + 1024; }

static int64_t comment_size_hint(const uint8_t* arc, const SegmentView& s);

// Device-side analysis of block payloads that are already in HBM: SHA-1 and the level-5 gap statistics (zpq_prep.cu).
// sha1 (20 bytes per block) and gaps (2 ints per block) are host arrays; either may be null.
static int analyze_dev(zpq_ctx* ctx, int nblk, const uint8_t* const* d_data, const uint64_t* len, uint8_t* sha1, int32_t* gaps) {
  if (nblk <= 0 || (!sha1 && !gaps)) return ZPQ_OK;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  cudaStream_t st = ctx->stream;
  const uint64_t o_ptr = 0, o_len = align_up((uint64_t)nblk * 8, 256), o_sha = o_len + align_up((uint64_t)nblk * 8, 256),
                 o_gap = o_sha + align_up((uint64_t)nblk * 20, 256), total = o_gap + align_up((uint64_t)nblk * 8, 256);
  CK(ctx->d_prep.reserve(total), "analysis cudaMalloc");
  uint8_t* dp = (uint8_t*)ctx->d_prep.p;
  CK(cudaMemcpyAsync(dp + o_ptr, d_data, (size_t)nblk * 8, cudaMemcpyHostToDevice, st), "H2D pointers");
  CK(cudaMemcpyAsync(dp + o_len, len, (size_t)nblk * 8, cudaMemcpyHostToDevice, st), "H2D lengths");
  CK(cudaEventRecord(ctx->ev[0], st), "event");
  if (sha1) {
    CK(zpq_launch_sha1((const uint8_t* const*)(dp + o_ptr), (const uint64_t*)(dp + o_len), dp + o_sha, nblk, st), "zpq_sha1_kernel launch");
    CK(cudaMemcpyAsync(sha1, dp + o_sha, (size_t)nblk * 20, cudaMemcpyDeviceToHost, st), "D2H sha1");
    ++ctx->launches;
  }
  if (gaps) {
    CK(zpq_launch_gaps((const uint8_t* const*)(dp + o_ptr), (const uint64_t*)(dp + o_len), (int32_t*)(dp + o_gap), nblk, ctx->sm_count, st), "zpq_gaps_kernel launch");
    CK(cudaMemcpyAsync(gaps, dp + o_gap, (size_t)nblk * 8, cudaMemcpyDeviceToHost, st), "D2H gaps");
    ++ctx->launches;
  }
  CK(cudaEventRecord(ctx->ev[1], st), "event");
  CK(cudaStreamSynchronize(st), "analysis");
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
  ctx->ms_prep = ms;
  return ZPQ_OK;
}

// Batched libzpaq::compressBlock (libzpaq.cpp:7543-7731), context-model family only.  `dev`: data[] and out[] are device
// pointers (blocks scattered into HBM over NCCL); otherwise host pointers, staged through device buffers here.  Either
// way everything per-byte runs on the device: SHA-1 and the level-5 gap statistics (zpq_prep.cu), the coding kernels, the
// framing bytes around the coded data (a scatter kernel).  The host only turns (method, length, gaps) into header bytes.
static int compress_core(zpq_ctx* ctx, int nblk, const char* method, const uint8_t* const* data,
                         const uint64_t* len, const char* const* filenames, const char* const* comments,
                         int dosha1, uint8_t* const* out, const uint64_t* out_cap, uint64_t* out_len,
                         int32_t* status, bool dev) {
  if (!ctx || nblk < 0 || !method || (nblk && (!data || !len || !out || !out_cap || !out_len))) return ZPQ_E_ARG;
  if (nblk == 0) return ZPQ_OK;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  cudaStream_t stream = ctx->stream;
  const auto t_begin = std::chrono::steady_clock::now();
  std::vector<int> st(nblk, ZPQ_OK);
  for (int i = 0; i < nblk; ++i) if (len[i] > 0x7fffffffull - 4096) st[i] = ZPQ_E_ARG;

  // ---- payloads in HBM ----
  std::vector<const uint8_t*> d_data(nblk, nullptr);
  if (dev) for (int i = 0; i < nblk; ++i) d_data[i] = data[i];
  else {
    uint64_t tot = 0;
    std::vector<uint64_t> off(nblk);
    for (int i = 0; i < nblk; ++i) { off[i] = tot; if (st[i] == ZPQ_OK) tot = align_up(tot + len[i] + 4, 16); }
    CK(ctx->d_stage_in.reserve(tot + 16), "input staging cudaMalloc");
    for (int i = 0; i < nblk; ++i) {
      if (st[i] != ZPQ_OK) continue;
      d_data[i] = (const uint8_t*)ctx->d_stage_in.p + off[i];
      if (len[i]) CK(cudaMemcpyAsync((void*)d_data[i], data[i], len[i], cudaMemcpyHostToDevice, stream), "H2D input");
    }
  }
  // ---- per-block analysis on the device: SHA-1, gap statistics of level 5 ----
  const bool need_gaps = method_needs_gaps(method);
  std::vector<uint8_t> sha((size_t)nblk * 20, 0);
  std::vector<int32_t> gaps((size_t)nblk * 2, 0);
  {
    std::vector<uint64_t> alen(nblk);
    for (int i = 0; i < nblk; ++i) alen[i] = st[i] == ZPQ_OK ? len[i] : 0;
    std::vector<const uint8_t*> aptr(d_data);
    for (int i = 0; i < nblk; ++i) if (!aptr[i]) aptr[i] = (const uint8_t*)ctx->d_tables;      // any valid address for an empty message
    const int rc = analyze_dev(ctx, nblk, aptr.data(), alen.data(), dosha1 ? sha.data() : nullptr, need_gaps ? gaps.data() : nullptr);
    if (rc != ZPQ_OK) return rc;
  }
  // ---- host: header bytes, pre-processing kind, model prefix and block prologue from (method, length, gaps) ----
  std::vector<std::vector<uint8_t>> hdrs(nblk), pro(nblk), prefix(nblk);
  std::vector<int> prep(nblk, 0);
  std::vector<std::array<int, 7>> lzn(nblk);      // makeConfig's $1..$7 of every block (LZ77 parameters)
  for (int i = 0; i < nblk; ++i) {
    if (st[i] != ZPQ_OK) continue;
    std::string e;
    const int g2[2] = {gaps[2 * i], gaps[2 * i + 1]};
    int sc = 0, lz5[5] = {0, 0, 0, 0, 0};
    st[i] = plan_block(method, len[i], g2, &hdrs[i], &prep[i], &prefix[i], &e, &sc, lz5);
    if (st[i] != ZPQ_OK) { if (ctx->err.empty() || st[i] == ZPQ_E_METHOD) ctx->err = e; continue; }
    lzn[i] = {sc, prep[i], lz5[0], lz5[1], lz5[2], lz5[3], lz5[4]};
    if (((prep[i] & 3) == 1 || (prep[i] & 3) == 2) && lz5[1] != 0) {
      // LZ77 with a second, longer context order probes bytes past the end of the block in the reference
      // (libzpaq.cpp:6700, no bound on in[i+3]): not reproducible, and no digit method selects it
      st[i] = ZPQ_E_METHOD; ctx->err = "LZ77 with a second context order ($4 > 0) is not built";
      continue;
    }
    char cs[32];
    snprintf(cs, sizeof(cs), "%llu", (unsigned long long)len[i]);
    std::string comment = cs;                                   // libzpaq.cpp:7702-7703
    if (comments && comments[i]) { comment += " "; comment += comments[i]; }
    write_block_prologue(&pro[i], hdrs[i].data(), (int)hdrs[i].size(), filenames ? filenames[i] : nullptr, comment.c_str());
    const uint64_t tail = 4 + (dosha1 ? 21 : 1) + 1;
    if (out_cap[i] < pro[i].size() + tail + 8) st[i] = ZPQ_E_OUTPUT_FULL;
  }
  // ---- pre-processing on the device (LZBuffer level 3 and e8e9(), libzpaq.cpp:6450-6459, 6630-6658): zpq_bwt.cu ----
  std::vector<const uint8_t*> d_model_in(d_data);                // what goes through the model after the prefix
  std::vector<uint64_t> model_len(len, len + nblk);
  ctx->ms_xform = 0; ctx->xform_bytes = 0;
  {
    uint64_t tot = 0;
    std::vector<uint64_t> off_e8(nblk, 0), off_bwt(nblk, 0), off_lz(nblk, 0);
    bool any = false;
    for (int i = 0; i < nblk; ++i) {
      if (st[i] != ZPQ_OK || prep[i] == 0) continue;
      any = true;
      if (prep[i] & 4) { off_e8[i] = tot; tot = align_up(tot + len[i] + 16, 256); }
      if ((prep[i] & 3) == 3) { off_bwt[i] = tot; tot = align_up(tot + len[i] + 5 + 16, 256); }
      if ((prep[i] & 3) == 1 || (prep[i] & 3) == 2) { off_lz[i] = tot; tot = align_up(tot + len[i] + len[i] / 8 + 1024 + 16, 256); }
    }
    if (any) {
      CK(ctx->d_xform.reserve(tot + 256), "pre-processing cudaMalloc");
      uint8_t* dx = (uint8_t*)ctx->d_xform.p;
      CK(cudaEventRecord(ctx->ev[0], stream), "event");
      // E8E9 into a working copy
      {
        std::vector<const uint8_t*> ein; std::vector<uint8_t*> eout; std::vector<uint64_t> elen;
        for (int i = 0; i < nblk; ++i) {
          if (st[i] != ZPQ_OK || !(prep[i] & 4)) continue;
          if (len[i]) CK(cudaMemcpyAsync(dx + off_e8[i], d_data[i], len[i], cudaMemcpyDeviceToDevice, stream), "E8E9 working copy");
          ein.push_back(d_data[i]); eout.push_back(dx + off_e8[i]); elen.push_back(len[i]);
          d_model_in[i] = dx + off_e8[i];
        }
        for (size_t k = 0; k < ein.size(); k += 60000) {
          const int cnt = (int)std::min<size_t>(60000, ein.size() - k);
          CK(ctx->d_xws.reserve((uint64_t)cnt * 24 + 1024), "pre-processing cudaMalloc");
          CK(zpq_e8e9_forward(ein.data() + k, eout.data() + k, elen.data() + k, cnt, ctx->d_xws.p, ctx->sm_count, stream, &ctx->launches), "zpq_e8e9_kernel launch");
        }
      }
      // BWT in sub-batches of at most 256 MiB (the sorter keeps ~51 bytes of state per payload byte)
      {
        std::vector<int> ids;
        for (int i = 0; i < nblk; ++i) if (st[i] == ZPQ_OK && (prep[i] & 3) == 3) ids.push_back(i);
        size_t k = 0;
        while (k < ids.size()) {
          uint64_t sum = 0; size_t e = k;
          while (e < ids.size() && e - k < 60000 && (e == k || sum + len[ids[e]] <= ((uint64_t)1 << 28))) sum += len[ids[e++]];
          std::vector<const uint8_t*> bin; std::vector<uint8_t*> bout; std::vector<uint64_t> blen;
          for (size_t j = k; j < e; ++j) {
            const int i = ids[j];
            bin.push_back(d_model_in[i]); bout.push_back(dx + off_bwt[i]); blen.push_back(len[i]);
          }
          const uint64_t need = zpq_bwt_workspace(sum, (int)(e - k));
          if (ctx->d_xws.reserve(need) != cudaSuccess) {
            cudaGetLastError();
            for (size_t j = k; j < e; ++j) st[ids[j]] = ZPQ_E_NOMEM;
            ctx->err = "Out of memory";
          }
          else {
            CK(zpq_bwt_forward(bin.data(), blen.data(), bout.data(), (int)(e - k), ctx->d_xws.p, ctx->d_xws.cap, ctx->sm_count, stream, &ctx->launches, &ctx->xform_bytes), "BWT kernels");
            for (size_t j = k; j < e; ++j) { const int i = ids[j]; d_model_in[i] = dx + off_bwt[i]; model_len[i] = len[i] + 5; }
          }
          k = e;
        }
      }
      // LZ77 (LZBuffer levels 1 / 2): one thread per block (zpq_lz77.cu); blocks with the same numbers go together, the
      // suffix-array variant sorts its sub-batch with the BWT's sorter first
      {
        std::vector<char> done(nblk, 0);
        for (int i0 = 0; i0 < nblk; ++i0) {
          if (done[i0] || st[i0] != ZPQ_OK || ((prep[i0] & 3) != 1 && (prep[i0] & 3) != 2)) continue;
          std::vector<int> ids;
          for (int i = i0; i < nblk; ++i)
            if (!done[i] && st[i] == ZPQ_OK && lzn[i] == lzn[i0]) { ids.push_back(i); done[i] = 1; }
          const int* numbers = lzn[i0].data();
          const bool use_sa = zpq_lz77_uses_sa(numbers) != 0;
          const uint64_t tb = zpq_lz77_table_bytes(numbers);
          size_t k = 0;
          while (k < ids.size()) {
            uint64_t sum = 0; size_t e = k;
            while (e < ids.size() && e - k < 60000 && (e == k || (sum + len[ids[e]] <= ((uint64_t)1 << 28) && (e - k + 1) * tb <= ((uint64_t)8 << 30)))) sum += len[ids[e++]];
            const int cnt = (int)(e - k);
            std::vector<const uint8_t*> lin; std::vector<uint8_t*> lout; std::vector<uint64_t> llen, lcap, olen(cnt, 0); std::vector<int> lst(cnt, 0);
            for (size_t j = k; j < e; ++j) {
              const int i = ids[j];
              lin.push_back(d_model_in[i]); lout.push_back(dx + off_lz[i]); llen.push_back(len[i]); lcap.push_back(len[i] + len[i] / 8 + 1024);
            }
            const uint32_t *d_sa = nullptr, *d_rank = nullptr, *d_off = nullptr;
            std::vector<uint32_t> base(cnt, 0);
            bool ok = ctx->d_lz.reserve(tb * (uint64_t)cnt + zpq_lz77_meta_bytes(cnt)) == cudaSuccess;
            if (ok && use_sa) {
              ok = ctx->d_xws.reserve(zpq_bwt_workspace(sum, cnt)) == cudaSuccess;
              if (ok) {
                CK(zpq_suffix_arrays(lin.data(), llen.data(), cnt, ctx->d_xws.p, ctx->d_xws.cap, ctx->sm_count, stream, &ctx->launches, &ctx->xform_bytes,
                                     &d_sa, &d_rank, &d_off), "suffix sort kernels");
                for (int j = 1; j < cnt; ++j) base[j] = base[j - 1] + (uint32_t)llen[j - 1];
              }
            }
            if (!ok) {
              cudaGetLastError();
              for (size_t j = k; j < e; ++j) st[ids[j]] = ZPQ_E_NOMEM;
              ctx->err = "Out of memory";
            } else {
              CK(zpq_lz77_forward(lin.data(), llen.data(), lout.data(), lcap.data(), cnt, numbers, ctx->d_lz.p, d_sa, d_rank, base.data(),
                                  (uint8_t*)ctx->d_lz.p + tb * (uint64_t)cnt, olen.data(), lst.data(), stream, &ctx->launches), "zpq_lz77_kernel launch");
              for (size_t j = k; j < e; ++j) {
                const int i = ids[j];
                if (lst[j - k]) { st[i] = ZPQ_E_NOMEM; continue; }
                d_model_in[i] = dx + off_lz[i]; model_len[i] = olen[j - k];
              }
            }
            k = e;
          }
        }
      }
      CK(cudaEventRecord(ctx->ev[1], stream), "event");
      CK(cudaStreamSynchronize(stream), "pre-processing");
      float ms = 0;
      cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
      ctx->ms_xform = ms;
    }
  }
  // ---- output blocks in HBM ----
  std::vector<uint8_t*> d_out(nblk, nullptr);
  if (dev) for (int i = 0; i < nblk; ++i) d_out[i] = out[i];
  else {
    uint64_t tot = 0;
    std::vector<uint64_t> off(nblk);
    for (int i = 0; i < nblk; ++i) { off[i] = tot; if (st[i] == ZPQ_OK) tot = align_up(tot + out_cap[i], 16); }
    CK(ctx->d_stage_out.reserve(tot + 16), "output staging cudaMalloc");
    for (int i = 0; i < nblk; ++i) if (st[i] == ZPQ_OK) d_out[i] = (uint8_t*)ctx->d_stage_out.p + off[i];
  }
  // ---- the coding path ----
  std::vector<IJob> jobs;
  std::vector<int> idx;
  std::vector<uint64_t> stored_len(nblk, 0);          // body length of blocks without a model (stored in 64 KiB chunks)
  std::vector<uint8_t> spool; std::vector<uint64_t> s_src; std::vector<uint8_t*> s_dst; std::vector<uint32_t> s_nb;
  for (int i = 0; i < nblk; ++i) {
    if (st[i] != ZPQ_OK) continue;
    const uint64_t tail = 4 + (dosha1 ? 21 : 1) + 1;
    if (hdrs[i][6] == 0) {
      // no components: Encoder::compress without a model (libzpaq.cpp:2436-2445) -- the bytes that would go through the model
      // (prefix, then the possibly transformed payload) leave in chunks of 64 KiB, each behind its length (4 bytes, MSB first)
      const uint64_t P = prefix[i].size(), L = P + model_len[i];
      const uint64_t body = L + 4 * ((L + 65535) / 65536);
      if (out_cap[i] < pro[i].size() + body + tail) { st[i] = ZPQ_E_OUTPUT_FULL; continue; }
      uint8_t* dst = d_out[i] + pro[i].size();
      for (uint64_t at = 0; at < L; at += 65536) {
        const uint64_t cl = std::min<uint64_t>(65536, L - at);
        const size_t h0 = spool.size();
        spool.push_back((uint8_t)(cl >> 24)); spool.push_back((uint8_t)(cl >> 16)); spool.push_back((uint8_t)(cl >> 8)); spool.push_back((uint8_t)cl);
        uint64_t from_prefix = at < P ? std::min<uint64_t>(cl, P - at) : 0;
        spool.insert(spool.end(), prefix[i].begin() + (at < P ? at : P), prefix[i].begin() + (at < P ? at : P) + from_prefix);
        s_src.push_back(h0); s_dst.push_back(dst); s_nb.push_back((uint32_t)(4 + from_prefix));
        dst += 4 + from_prefix;
        const uint64_t from_data = cl - from_prefix;
        if (from_data) {
          const uint64_t doff = at + from_prefix - P;
          CK(cudaMemcpyAsync(dst, d_model_in[i] + doff, from_data, cudaMemcpyDeviceToDevice, stream), "stored chunk");
          dst += from_data;
        }
      }
      stored_len[i] = body;
      continue;
    }
    IJob J;
    J.hdr = hdrs[i].data(); J.hlen = (int)hdrs[i].size();
    J.in.push_back(Piece{prefix[i].data(), prefix[i].size()});      // PASS byte, or "1, len, PCOMP program" (libzpaq.cpp:2905-2915)
    J.in.push_back(Piece{d_model_in[i], model_len[i]});
    J.seg_len.assign(1, model_len[i] + prefix[i].size());
    J.out = d_out[i] + pro[i].size();
    J.out_cap = out_cap[i] - pro[i].size() - tail;
    jobs.push_back(J); idx.push_back(i);
  }
  int rc = run_batch(ctx, jobs, false, false);
  const float ms_init = ctx->ms[0] + ctx->ms_prep + ctx->ms_xform, ms_code = ctx->ms[1];
  // ---- framing around the coded data (Compressor::endSegment / endBlock, libzpaq.cpp:2956-3002), written on the device ----
  {
    std::vector<uint8_t> pool; std::vector<uint64_t> src_off; std::vector<uint8_t*> dst; std::vector<uint32_t> nb;
    std::vector<uint64_t> body_len(nblk, 0); std::vector<char> have(nblk, 0);
    for (size_t k = 0; k < jobs.size(); ++k) {
      const int i = idx[k];
      st[i] = jobs[k].status;
      if (st[i] == ZPQ_OK) { body_len[i] = jobs[k].out_len; have[i] = 1; }
    }
    for (int i = 0; i < nblk; ++i) if (st[i] == ZPQ_OK && hdrs[i].size() > 6 && hdrs[i][6] == 0) { body_len[i] = stored_len[i]; have[i] = 1; }
    // chunk headers (and program bytes) of the stored blocks share the framing pool
    for (size_t r = 0; r < s_src.size(); ++r) { src_off.push_back(pool.size() + s_src[r]); dst.push_back(s_dst[r]); nb.push_back(s_nb[r]); }
    pool.insert(pool.end(), spool.begin(), spool.end());
    for (int i = 0; i < nblk; ++i) {
      if (!have[i] || st[i] != ZPQ_OK) continue;
      struct { uint64_t out_len; } jk = {body_len[i]};
      const auto* jobs_k = &jk;
      src_off.push_back(pool.size()); dst.push_back(d_out[i]); nb.push_back((uint32_t)pro[i].size());
      pool.insert(pool.end(), pro[i].begin(), pro[i].end());
      src_off.push_back(pool.size()); dst.push_back(d_out[i] + pro[i].size() + jobs_k->out_len);
      const size_t t0 = pool.size();
      pool.insert(pool.end(), 4, 0);
      if (dosha1) { pool.push_back(253); pool.insert(pool.end(), sha.begin() + (size_t)i * 20, sha.begin() + (size_t)i * 20 + 20); }
      else pool.push_back(254);
      pool.push_back(255);
      nb.push_back((uint32_t)(pool.size() - t0));
      out_len[i] = pro[i].size() + jobs_k->out_len + (pool.size() - t0);
    }
    const int nrec = (int)dst.size();
    if (nrec) {
      const uint64_t o_pool = 0, o_src = align_up(pool.size(), 256), o_dst = o_src + align_up((uint64_t)nrec * 8, 256),
                     o_nb = o_dst + align_up((uint64_t)nrec * 8, 256), total = o_nb + align_up((uint64_t)nrec * 4, 256);
      CK(ctx->d_prep.reserve(total), "framing cudaMalloc");
      uint8_t* dp = (uint8_t*)ctx->d_prep.p;
      CK(cudaMemcpyAsync(dp + o_pool, pool.data(), pool.size(), cudaMemcpyHostToDevice, stream), "H2D framing");
      CK(cudaMemcpyAsync(dp + o_src, src_off.data(), (size_t)nrec * 8, cudaMemcpyHostToDevice, stream), "H2D framing");
      CK(cudaMemcpyAsync(dp + o_dst, dst.data(), (size_t)nrec * 8, cudaMemcpyHostToDevice, stream), "H2D framing");
      CK(cudaMemcpyAsync(dp + o_nb, nb.data(), (size_t)nrec * 4, cudaMemcpyHostToDevice, stream), "H2D framing");
      CK(zpq_launch_scatter(dp + o_pool, (const uint64_t*)(dp + o_src), (uint8_t* const*)(dp + o_dst), (const uint32_t*)(dp + o_nb), nrec, stream), "zpq_scatter_kernel launch");
      ++ctx->launches;
    }
    if (!dev)
      for (int i = 0; i < nblk; ++i)
        if (st[i] == ZPQ_OK) CK(cudaMemcpyAsync(out[i], d_out[i], out_len[i], cudaMemcpyDeviceToHost, stream), "D2H blocks");
    CK(cudaStreamSynchronize(stream), "framing");
  }
  ctx->ms[0] = ms_init; ctx->ms[1] = ms_code;
  ctx->ms[2] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  int first_bad = ZPQ_OK;
  for (int i = 0; i < nblk; ++i) {
    if (status) status[i] = st[i];
    if (st[i] != ZPQ_OK && first_bad == ZPQ_OK) first_bad = st[i];
  }
  return rc != ZPQ_OK ? rc : first_bad;
}

int zpq_compress_blocks(zpq_ctx* ctx, int nblk, const char* method, const uint8_t* const* data,
                        const uint64_t* len, const char* const* filenames, const char* const* comments,
                        int dosha1, uint8_t* const* out, const uint64_t* out_cap, uint64_t* out_len,
                        int32_t* status) {
  return compress_core(ctx, nblk, method, data, len, filenames, comments, dosha1, out, out_cap, out_len, status, false);
}

int zpq_compress_blocks_dev(zpq_ctx* ctx, int nblk, const char* method, const uint8_t* const* d_data,
                            const uint64_t* len, const char* const* filenames, const char* const* comments,
                            int dosha1, uint8_t* const* d_out, const uint64_t* out_cap, uint64_t* out_len,
                            int32_t* status) {
  return compress_core(ctx, nblk, method, d_data, len, filenames, comments, dosha1, d_out, out_cap, out_len, status, true);
}

int zpq_analyze_dev(zpq_ctx* ctx, int nblk, const uint8_t* const* d_data, const uint64_t* len, uint8_t* sha1, int32_t* gaps) {
  if (!ctx || nblk < 0 || (nblk && (!d_data || !len))) return ZPQ_E_ARG;
  return analyze_dev(ctx, nblk, d_data, len, sha1, gaps);
}

// Framing scan for callers that bind at the hot-path level (block sharding): one record per segment.
int zpq_scan_archive(const uint8_t* arc, uint64_t arc_len, zpq_segment* segs, int cap, int* nsegs, int* nblocks) {
  if (!arc || !nsegs) return ZPQ_E_ARG;
  std::vector<BlockView> bv; std::string e;
  const int rc = scan_archive(arc, arc_len, &bv, &e);
  if (rc != ZPQ_OK) return rc;
  int n = 0;
  for (size_t b = 0; b < bv.size(); ++b)
    for (size_t k = 0; k < bv[b].segs.size(); ++k, ++n) {
      if (!segs || n >= cap) continue;
      zpq_segment& S = segs[n];
      const SegmentView& V = bv[b].segs[k];
      S.block = (uint32_t)b; S.segment = (uint32_t)k;
      S.hdr_off = bv[b].hdr_off; S.hlen = (uint32_t)bv[b].hlen; S.components = arc[bv[b].hdr_off + 6];
      S.coded_off = V.coded_off; S.coded_len = V.coded_len;
      S.has_sha1 = V.has_sha1 ? 1 : 0; memcpy(S.sha1, V.sha1, 20);
      S.size_hint = comment_size_hint(arc, V);
    }
  *nsegs = n;
  if (nblocks) *nblocks = (int)bv.size();
  return (segs && n > cap) ? ZPQ_E_OUTPUT_FULL : ZPQ_OK;
}

// Batched libzpaq::decompress (libzpaq.cpp:2376-2389) over an in-memory archive.
// The block framing does not store decoded sizes.  compressBlock writes the
// size as the decimal prefix of the segment comment (libzpaq.cpp:7702), so
// when every segment carries one the payloads are decoded straight into
// their final positions; otherwise blocks go through host staging buffers
// whose capacity grows until the decode fits.
static int64_t comment_size_hint(const uint8_t* arc, const SegmentView& s) {
  if (s.comment_len == 0 || s.comment_len > 40) return -1;
  uint64_t v = 0; uint64_t k = 0;
  const uint8_t* c = arc + s.comment_off;
  while (k < s.comment_len && c[k] >= '0' && c[k] <= '9' && k < 12) v = v * 10 + (c[k++] - '0');
  if (k == 0 || (k < s.comment_len && c[k] != ' ')) return -1;
  // the comment is untrusted: a size that no coder could have produced from this many coded bytes (the arithmetic coder
  // spends at least ~1/700 byte per input byte on fully predicted data; 2^16 x is far beyond that), or beyond the 2^31
  // block limit, is not a hint -- the block then goes through the grow-until-it-fits staging path
  if (v > ((uint64_t)1 << 31) - 4096 || v > (s.coded_len + 64) * 65536ull) return -1;
  return (int64_t)v;
}

int zpq_decompress_archive(zpq_ctx* ctx, const uint8_t* arc, uint64_t arc_len, uint8_t* out, uint64_t out_cap,
                           uint64_t* out_len, int* nblocks, int verify_sha1) {
  if (!ctx || !arc || !out_len) return ZPQ_E_ARG;
  *out_len = 0;
  std::vector<BlockView> bv; std::string e;
  int rc = scan_archive(arc, arc_len, &bv, &e);
  if (rc != ZPQ_OK) { ctx->err = e; return rc; }
  if (nblocks) *nblocks = (int)bv.size();
  const size_t nb = bv.size();
  if (nb == 0) return ZPQ_OK;

  std::vector<IJob> jobs(nb);
  std::vector<int64_t> hint(nb, 0);
  bool all_hinted = true;
  for (size_t i = 0; i < nb; ++i) {
    IJob& J = jobs[i];
    J.hdr = arc + bv[i].hdr_off; J.hlen = bv[i].hlen;
    for (const SegmentView& s : bv[i].segs) {
      J.in.push_back(Piece{arc + s.coded_off, s.coded_len});
      J.seg_len.push_back(s.coded_len);
      const int64_t h = comment_size_hint(arc, s);
      if (h < 0 || hint[i] < 0) hint[i] = -1; else hint[i] += h;
    }
    if (bv[i].segs.empty()) hint[i] = -1;
    if (hint[i] < 0) all_hinted = false;
  }
  int first_bad = ZPQ_OK;
  auto note = [&](int code) { if (code != ZPQ_OK && first_bad == ZPQ_OK) first_bad = code; };

  std::vector<std::vector<uint8_t>> tmp;
  std::vector<uint64_t> pos(nb, 0);
  // unmodelled blocks (n = 0, level 2) have nothing to decode: their "model output" is the stored chunks back to back
  // (4-byte big-endian length + bytes, zero length ends the segment; Decoder::decompress libzpaq.cpp:2138-2153).  They
  // are unpacked here on the host and only meet the other blocks again at the post-processing step.
  std::vector<char> stored(nb, 0);
  bool any_stored = false;
  for (size_t i = 0; i < nb; ++i) if (jobs[i].hlen > 6 && jobs[i].hdr[6] == 0) { stored[i] = 1; any_stored = true; }
  // a header that announces a post-processor machine (ph / pm: BWT and LZ77 blocks) means the model output is a PCOMP program
  // plus transformed bytes, never the payload itself: such an archive goes through staging at once instead of finding out
  // by a wasted decode into the caller's buffer
  bool any_pcomp = false;
  for (size_t i = 0; i < nb; ++i) if (jobs[i].hlen > 6 && (jobs[i].hdr[4] || jobs[i].hdr[5])) any_pcomp = true;
  bool direct = all_hinted && !any_stored && !any_pcomp;
  if (direct) {
    uint64_t p = 0;
    for (size_t i = 0; i < nb; ++i) { pos[i] = p; p += (uint64_t)hint[i]; }
    if (p > out_cap || !out) direct = false;
  }
  if (direct) {
    for (size_t i = 0; i < nb; ++i) {
      jobs[i].out = out + pos[i]; jobs[i].out_cap = (uint64_t)hint[i]; jobs[i].out_skip = 1;   // drop the PASS flag byte
    }
    rc = run_batch(ctx, jobs, true, false);
    if (rc != ZPQ_OK) return rc;
    for (size_t i = 0; i < nb && direct; ++i)
      if (jobs[i].status == ZPQ_E_OUTPUT_FULL || (jobs[i].status == ZPQ_OK && (jobs[i].out_len != (uint64_t)hint[i] + 1 || jobs[i].first_byte != 0)))
        direct = false;     // a comment lied about the size, or the block carries a PCOMP program: redo through staging
  }
  // blocks whose model output starts with 1 (PROG) are post-processed on the host after staging (zpq_hcomp.cpp postprocess_block)
  std::vector<std::vector<uint8_t>> post(nb);
  std::vector<std::vector<uint64_t>> post_seg(nb);
  std::vector<char> is_prog(nb, 0);
  std::vector<int> post_rc(nb, ZPQ_OK);
  if (!direct) {
    tmp.resize(nb);
    std::vector<size_t> pending;
    for (size_t i = 0; i < nb; ++i) {
      if (!stored[i]) { pending.push_back(i); continue; }
      IJob& J = jobs[i];
      J.status = ZPQ_OK; J.seg_out.assign(bv[i].segs.size(), 0);
      for (size_t sg = 0; sg < bv[i].segs.size() && J.status == ZPQ_OK; ++sg) {
        const uint8_t* a = arc + bv[i].segs[sg].coded_off; const uint64_t n = bv[i].segs[sg].coded_len;
        uint64_t e = 0;
        for (;;) {
          if (e + 4 > n) { J.status = ZPQ_E_EOF; break; }
          const uint64_t len = (uint64_t)a[e] << 24 | (uint64_t)a[e + 1] << 16 | (uint64_t)a[e + 2] << 8 | a[e + 3];
          e += 4;
          if (len == 0) break;
          if (e + len > n) { J.status = ZPQ_E_EOF; break; }
          tmp[i].insert(tmp[i].end(), a + e, a + e + len);
          J.seg_out[sg] += len;
          e += len;
        }
      }
      J.out_len = tmp[i].size();
      J.first_byte = tmp[i].empty() ? -1 : tmp[i][0];
    }
    std::vector<uint64_t> cap(nb);
    // size hint + flag byte + room for a PCOMP program (<= 64 KiB + 3) and a transform's trailer (BWT: 5 bytes): a block with
    // a program must not be decoded twice only because its program did not fit the payload's size
    for (size_t i = 0; i < nb; ++i) cap[i] = hint[i] >= 0 ? (uint64_t)hint[i] + 1 + 66000 : (uint64_t)4 << 20;
    while (!pending.empty()) {
      std::vector<IJob> sub;
      for (size_t i : pending) {
        try { tmp[i].resize(cap[i]); }
        catch (const std::bad_alloc&) { jobs[i].status = ZPQ_E_NOMEM; tmp[i].clear(); tmp[i].resize(16); cap[i] = 16; }
        IJob J = jobs[i];
        J.out = tmp[i].data(); J.out_cap = cap[i]; J.out_skip = 0;
        sub.push_back(J);
      }
      rc = run_batch(ctx, sub, true, false);
      if (rc != ZPQ_OK) return rc;
      std::vector<size_t> again;
      for (size_t k = 0; k < pending.size(); ++k) {
        const size_t i = pending[k];
        if (jobs[i].status == ZPQ_E_NOMEM) continue;                   // its staging buffer could not be allocated above
        jobs[i].status = sub[k].status; jobs[i].out_len = sub[k].out_len; jobs[i].seg_out = sub[k].seg_out;
        jobs[i].first_byte = tmp[i].empty() ? -1 : tmp[i][0];
        if (sub[k].status == ZPQ_E_OUTPUT_FULL && cap[i] < ((uint64_t)1 << 31)) {
          cap[i] = std::min<uint64_t>(std::max(cap[i] * 8, sub[k].out_len), (uint64_t)1 << 31);
          again.push_back(i);
        }
      }
      pending.swap(again);
    }
    {
      std::vector<size_t> progs;
      for (size_t i = 0; i < nb; ++i)
        if (jobs[i].status == ZPQ_OK && jobs[i].out_len >= 1 && tmp[i][0] == 1) { is_prog[i] = 1; progs.push_back(i); }
      // Blocks whose program is one of the post-processors libzpaq's compressBlock (and this library) writes are post-processed
      // on the device: the BWT inverse as an LF mapping by stable radix sort + list ranking (zpq_bwt.cu), the LZ77 decoders and
      // the E8E9 inverse with one thread per block (zpq_lz77.cu).  Every other program, a block without a usable size hint,
      // and any stream those kernels refuse run in the host ZPAQL machine below.
      {
        struct Rec { size_t i; int kind; size_t skip; uint64_t m; uint64_t n; uint32_t idx; int min_match; };
        std::vector<Rec> recs;
        const char* he = getenv("ZPQ_HOST_IBWT");                          // A/B switch: 1 = every program in the host machine
        const bool host_only = he && atoi(he) != 0;
        for (size_t i : progs) {
          if (host_only) break;
          const IJob& J = jobs[i];
          const int ph = J.hdr[4], pm = J.hdr[5];
          if (J.seg_out.size() != 1 || J.out_len < 3) continue;
          const size_t plen = 3 + (size_t)tmp[i][1] + 256 * (size_t)tmp[i][2];
          if (J.out_len < plen) continue;
          int cand[4] = {0, 0, 0, 0}, nc = 0, sc = 0;
          if (ph == pm && ph >= 20 && ph <= 31) { cand[0] = 3; cand[1] = 7; nc = 2; sc = ph - 20; }
          else if (ph == 0 && pm >= 20 && pm <= 31) { cand[0] = 1; cand[1] = 5; cand[2] = 2; cand[3] = 6; nc = 4; sc = pm - 20; }
          else if (ph == 0 && pm == 0) { cand[0] = 4; nc = 1; }
          for (int c = 0; c < nc; ++c) {
            const int kind = cand[c];
            int mm = 0;
            std::vector<uint8_t> code, pre;
            if ((kind & 3) == 2) {
              // the minimum match length is an immediate of the program: find where, read it, rebuild the program with it
              std::vector<uint8_t> c1, c2;
              if (pcomp_program(kind, sc, 1, &c1) != ZPQ_OK || pcomp_program(kind, sc, 2, &c2) != ZPQ_OK || c1.size() != c2.size()) continue;
              size_t at = 0;
              while (at < c1.size() && c1[at] == c2[at]) ++at;
              if (at >= c1.size() || 3 + at >= plen) continue;
              mm = tmp[i][3 + at];
              if (mm < 1) continue;
            }
            if (pcomp_program(kind, sc, mm, &code) != ZPQ_OK) continue;
            pcomp_prefix(code, &pre);
            if (pre.size() != plen || memcmp(tmp[i].data(), pre.data(), plen) != 0) continue;
            Rec r; r.i = i; r.kind = kind; r.skip = plen; r.m = J.out_len - plen; r.n = 0; r.idx = 0; r.min_match = mm;
            if ((kind & 3) == 3) {
              if (r.m < 6) break;
              r.n = r.m - 5;
              if (r.n + 5 > ((uint64_t)1 << pm)) break;
              const uint8_t* t = tmp[i].data() + plen + r.n + 1;
              r.idx = (uint32_t)t[0] | (uint32_t)t[1] << 8 | (uint32_t)t[2] << 16 | (uint32_t)t[3] << 24;
              if (r.idx < 1 || r.idx > r.n) break;
            }
            else if (kind == 4) r.n = r.m;
            else {
              if (hint[i] < 0 || (uint64_t)hint[i] + 4096 > ((uint64_t)1 << pm)) break;     // the decoders need the size up front
              r.n = (uint64_t)hint[i];
            }
            recs.push_back(r);
            break;
          }
        }
        // ---- BWT blocks (3, 7): inverse on the device, E8E9 undone in place for 7 ----
        size_t k = 0;
        std::vector<Rec> bw, lz;
        for (const Rec& r : recs) ((r.kind & 3) == 3 ? bw : lz).push_back(r);
        while (k < bw.size()) {
          uint64_t sum = 0; size_t e2 = k;
          while (e2 < bw.size() && e2 - k < 60000 && (e2 == k || sum + bw[e2].n + 1 <= ((uint64_t)1 << 28))) sum += bw[e2++].n + 1;
          const int cnt = (int)(e2 - k);
          uint64_t bytes = 0;
          std::vector<uint64_t> o_in(cnt), o_out(cnt), bn(cnt); std::vector<uint32_t> bidx(cnt);
          for (size_t j = k; j < e2; ++j) {
            o_in[j - k] = bytes; bytes = align_up(bytes + bw[j].n + 5 + 16, 256); o_out[j - k] = bytes; bytes = align_up(bytes + bw[j].n + 16, 256);
            bn[j - k] = bw[j].n; bidx[j - k] = bw[j].idx;
          }
          const uint64_t need = std::max<uint64_t>(zpq_ibwt_workspace(sum, cnt), zpq_unlz_meta_bytes(cnt));
          if (sum < ((uint64_t)1 << 31) && ctx->d_xform.reserve(bytes + 256) == cudaSuccess && ctx->d_xws.reserve(need) == cudaSuccess) {
            uint8_t* dx = (uint8_t*)ctx->d_xform.p;
            std::vector<const uint8_t*> pin; std::vector<uint8_t*> pout;
            for (size_t j = k; j < e2; ++j) {
              const size_t i = bw[j].i;
              CK(cudaMemcpyAsync(dx + o_in[j - k], tmp[i].data() + bw[j].skip, bw[j].n + 5, cudaMemcpyHostToDevice, ctx->stream), "H2D BWT rows");
              pin.push_back(dx + o_in[j - k]); pout.push_back(dx + o_out[j - k]);
            }
            CK(zpq_bwt_inverse(pin.data(), bn.data(), bidx.data(), pout.data(), cnt, ctx->d_xws.p, ctx->d_xws.cap,
                               ctx->sm_count, ctx->stream, &ctx->launches), "inverse BWT kernels");
            // E8E9 + BWT blocks: the operands are made relative again where they lie
            {
              std::vector<const uint8_t*> ein; std::vector<uint8_t*> eout; std::vector<uint64_t> elen;
              for (size_t j = k; j < e2; ++j) if (bw[j].kind == 7) { ein.push_back(pout[j - k]); eout.push_back(pout[j - k]); elen.push_back(bw[j].n); }
              if (!ein.empty()) {
                std::vector<uint64_t> ol(ein.size()); std::vector<int> os(ein.size());
                CK(zpq_unlz_forward(ein.data(), elen.data(), eout.data(), elen.data(), (int)ein.size(), 4, 1, 0, 0, ctx->d_xws.p, ol.data(), os.data(),
                                    ctx->stream, &ctx->launches), "zpq_unlz_kernel launch");
              }
            }
            for (size_t j = k; j < e2; ++j) {
              const size_t i = bw[j].i;
              post[i].resize(bw[j].n);
              CK(cudaMemcpyAsync(post[i].data(), pout[j - k], bw[j].n, cudaMemcpyDeviceToHost, ctx->stream), "D2H inverse BWT");
              post_seg[i].assign(1, bw[j].n);
            }
            CK(cudaStreamSynchronize(ctx->stream), "inverse BWT");
            for (size_t j = k; j < e2; ++j) {
              const size_t i = bw[j].i;
              is_prog[i] = 2;                                             // done on the device
              if (bw[j].kind == 7 && bv[i].segs[0].has_sha1) {            // E8E9 undone by zpq_unlz_kernel: same cross-check as below
                uint8_t dg[20];
                sha1_of(post[i].data(), post[i].size(), dg);
                if (memcmp(dg, bv[i].segs[0].sha1, 20) != 0) { is_prog[i] = 1; post[i].clear(); post_seg[i].clear(); }
              }
            }
          }
          else cudaGetLastError();                                        // no room: these blocks take the host machine
          k = e2;
        }
        // ---- LZ77 decoders (1, 2, 5, 6) and E8E9 alone (4): groups with the same program parameters ----
        std::vector<char> done(lz.size(), 0);
        for (size_t a = 0; a < lz.size(); ++a) {
          if (done[a]) continue;
          std::vector<size_t> g;
          const int sc_a = jobs[lz[a].i].hdr[5] >= 20 ? jobs[lz[a].i].hdr[5] - 20 : 0;
          for (size_t b2 = a; b2 < lz.size(); ++b2) {
            const int sc_b = jobs[lz[b2].i].hdr[5] >= 20 ? jobs[lz[b2].i].hdr[5] - 20 : 0;
            if (!done[b2] && lz[b2].kind == lz[a].kind && lz[b2].min_match == lz[a].min_match && sc_b == sc_a) { g.push_back(b2); done[b2] = 1; }
          }
          for (size_t k2 = 0; k2 < g.size(); k2 += 4096) {
            const int cnt = (int)std::min<size_t>(4096, g.size() - k2);
            uint64_t bytes = 0;
            std::vector<uint64_t> o_in(cnt), o_out(cnt), ilen(cnt), ocap(cnt), ol(cnt, 0); std::vector<int> os(cnt, 1);
            for (int j = 0; j < cnt; ++j) {
              const Rec& r = lz[g[k2 + j]];
              o_in[j] = bytes; bytes = align_up(bytes + r.m + 16, 256); o_out[j] = bytes; bytes = align_up(bytes + r.n + 16, 256);
              ilen[j] = r.m; ocap[j] = r.n;
            }
            if (ctx->d_xform.reserve(bytes + 256) != cudaSuccess || ctx->d_xws.reserve(zpq_unlz_meta_bytes(cnt)) != cudaSuccess) { cudaGetLastError(); continue; }
            uint8_t* dx = (uint8_t*)ctx->d_xform.p;
            std::vector<const uint8_t*> pin; std::vector<uint8_t*> pout;
            for (int j = 0; j < cnt; ++j) {
              const Rec& r = lz[g[k2 + j]];
              if (r.m) CK(cudaMemcpyAsync(dx + o_in[j], tmp[r.i].data() + r.skip, r.m, cudaMemcpyHostToDevice, ctx->stream), "H2D LZ77 codes");
              pin.push_back(dx + o_in[j]); pout.push_back(dx + o_out[j]);
            }
            const int kind = lz[a].kind, base_kind = kind == 4 ? 4 : (kind & 3), e8 = kind >= 4 ? 1 : 0;
            CK(zpq_unlz_forward(pin.data(), ilen.data(), pout.data(), ocap.data(), cnt, base_kind, e8, sc_a > 4 ? sc_a - 4 : 0, lz[a].min_match,
                                ctx->d_xws.p, ol.data(), os.data(), ctx->stream, &ctx->launches), "zpq_unlz_kernel launch");
            for (int j = 0; j < cnt; ++j) {
              const Rec& r = lz[g[k2 + j]];
              if (os[j] != 0 || ol[j] != r.n) continue;                   // refused, or not the announced size: the host machine decides
              post[r.i].resize(r.n);
              if (r.n) CK(cudaMemcpyAsync(post[r.i].data(), pout[j], r.n, cudaMemcpyDeviceToHost, ctx->stream), "D2H LZ77 output");
              post_seg[r.i].assign(1, r.n);
              is_prog[r.i] = 2;
            }
            CK(cudaStreamSynchronize(ctx->stream), "LZ77 decode");
            // a block that carries its checksum is checked here: if the device decoder and the checksum disagree the host
            // machine gets the block (it then either agrees with the checksum or the archive really is damaged)
            for (int j = 0; j < cnt; ++j) {
              const Rec& r = lz[g[k2 + j]];
              if (is_prog[r.i] != 2 || !bv[r.i].segs[0].has_sha1) continue;
              uint8_t dg[20];
              sha1_of(post[r.i].data(), post[r.i].size(), dg);
              if (memcmp(dg, bv[r.i].segs[0].sha1, 20) != 0) { is_prog[r.i] = 1; post[r.i].clear(); post_seg[r.i].clear(); }
            }
          }
        }
        std::vector<size_t> rest;
        for (size_t i : progs) if (is_prog[i] != 2) rest.push_back(i);
        progs.swap(rest);
      }
      const unsigned nt = (unsigned)std::min<size_t>(progs.size(), std::max(1u, std::min(std::thread::hardware_concurrency(), 64u)));
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t]() {
          for (size_t k = t; k < progs.size(); k += nt) {
            const size_t i = progs[k];
            std::string pe;
            post_rc[i] = postprocess_block(jobs[i].hdr[4], jobs[i].hdr[5], tmp[i].data(), jobs[i].seg_out.data(), (int)jobs[i].seg_out.size(),
                                           &post[i], &post_seg[i], &pe);
          }
        });
      for (auto& t : th) t.join();
    }
    uint64_t p = 0;
    for (size_t i = 0; i < nb; ++i) {
      pos[i] = p;
      if (jobs[i].status != ZPQ_OK || jobs[i].out_len < 1 || post_rc[i] != ZPQ_OK) continue;
      const uint64_t n = is_prog[i] ? post[i].size() : jobs[i].out_len - 1;
      const uint8_t* src = is_prog[i] ? post[i].data() : tmp[i].data() + 1;
      if (p + n <= out_cap && out) { if (n) memcpy(out + p, src, n); }
      else note(ZPQ_E_OUTPUT_FULL);
      p += n;
    }
  }
  // per-block checks: status, post-processor flag, stored checksums
  std::vector<int> st(nb, ZPQ_OK);
  {
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    nt = std::min<unsigned>(nt, (unsigned)nb);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([&, t]() {
        for (size_t i = t; i < nb; i += nt) {
          const IJob& J = jobs[i];
          if (J.status != ZPQ_OK) { st[i] = J.status; continue; }
          if (J.out_len < 1) { st[i] = ZPQ_E_CORRUPT; continue; }
          if (post_rc[i] != ZPQ_OK) { st[i] = post_rc[i]; continue; }
          if (J.first_byte != 0 && !is_prog[i]) { st[i] = ZPQ_E_CORRUPT; continue; }   // "unknown post processing type"
          const uint64_t total_i = is_prog[i] ? post[i].size() : J.out_len - 1;
          if (!verify_sha1 || pos[i] + total_i > out_cap || !out) continue;
          uint64_t at = pos[i];
          for (size_t s = 0; s < J.seg_out.size(); ++s) {
            if (!is_prog[i] && s == 0 && J.seg_out[0] == 0) { st[i] = ZPQ_E_CORRUPT; break; }    // no room for the post-processor flag byte
            const uint64_t n = is_prog[i] ? post_seg[i][s] : J.seg_out[s] - (s == 0 ? 1 : 0);
            if (at + n > out_cap) { st[i] = ZPQ_E_CORRUPT; break; }
            if (bv[i].segs[s].has_sha1) {
              uint8_t d[20];
              sha1_of(out + at, n, d);
              if (memcmp(d, bv[i].segs[s].sha1, 20) != 0) st[i] = ZPQ_E_CHECKSUM;
            }
            at += n;
          }
        }
      });
    for (auto& t : th) t.join();
  }
  uint64_t total = 0;
  for (size_t i = 0; i < nb; ++i) {
    note(st[i]);
    if (st[i] == ZPQ_OK || st[i] == ZPQ_E_CHECKSUM) total = pos[i] + (is_prog[i] ? post[i].size() : jobs[i].out_len - 1);
  }
  *out_len = total;
  return first_bad;
}

}  // extern "C"
