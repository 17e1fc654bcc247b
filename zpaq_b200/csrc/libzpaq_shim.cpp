// libzpaq_shim.cpp -- link-compatible replacements of libzpaq::compressBlock, Decompresser::decompress and
// Decompresser::readSegmentEnd for the UNCHANGED zpaq.cpp.
//
// Compiled against the unchanged reference header (libzpaq.h, never copied into this repo; the build
// recipe oracle/Makefile adds -I/root/reference) and linked in front of the reference library object,
// whose own three functions are renamed by objcopy (compressBlock_cpu, zpq_ref_decompress,
// zpq_ref_readSegmentEnd).  zpaq.cpp's compress threads (zpaq.cpp:1918-1965) and decompress threads
// (zpaq.cpp:2796-2870) call these functions one block at a time; concurrent calls are coalesced into one
// GPU batch by a small submitter so that the device sees many blocks per launch.
//
//   * context-model methods (-method 4, 5, x0,0c...)  -> zpq_compress_blocks (CUDA, bit-identical bytes)
//   * methods that need LZ77/BWT/E8E9 pre-processing   -> the reference's own CPU code (out of scope of
//     the hot path)
//   * a context-model method without a usable GPU      -> libzpaq::error("CUDA error"): no CPU fallback
//     of the hot path exists in libzpq_b200.so.
//   * Decompresser::decompress(n): the first call of a segment pulls the segment's coded bytes from the
//     Reader (through Decoder::get, so buffered() keeps its meaning), decodes them with zpq_decode_blocks
//     and then serves n bytes per call through the object's own PostProcessor (PASS and PCOMP alike:
//     post-processing is not part of the hot path and stays the reference's code).  Stored blocks (no model)
//     keep the reference's CPU code; a MODELLED block the kernels cannot take raises error(), never a CPU decode.
#include "libzpaq.h"
#include "../../include/zpq.h"
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <chrono>
#include <map>
#include <string.h>
#include <stdlib.h>
#include <ctype.h>
#include <stdio.h>
#include <algorithm>

namespace libzpaq {

// the reference implementation under its objcopy-renamed symbol
void compressBlock_cpu(StringBuffer* in, Writer* out, const char* method, const char* filename, const char* comment, bool dosha1);

namespace {

// ZPQ_SHIM_LOG=1: one line per routed call on stderr (which blocks went to the GPU library, which stayed on the
// reference's CPU code) -- used by the tests to prove that the CUDA path was the one that ran
bool shim_log() { static const bool on = [] { const char* e = getenv("ZPQ_SHIM_LOG"); return e && atoi(e) != 0; }(); return on; }
#define SHIM_LOG(...) do { if (shim_log()) { fprintf(stderr, "[zpq shim] " __VA_ARGS__); fputc('\n', stderr); } } while (0)

struct Req {
  bool decode = false;
  // compress: one whole block
  const uint8_t* data = nullptr; uint64_t len = 0; std::string method; const char* filename = nullptr; const char* comment = nullptr; bool dosha1 = false;
  // decode: header as stored, the coded segments of the block so far (each with its four zeros), output capacity
  const uint8_t* hdr = nullptr; int hlen = 0; const std::vector<uint64_t>* seg_len = nullptr; std::vector<uint64_t> seg_out;
  std::vector<uint8_t> out; uint64_t out_len = 0; int status = 1; bool done = false;
};

// how long the submitter waits for more blocks before it launches (zpaq keeps at most `-threads` blocks in flight, and a
// launch is bound by the latency of one block, so waiting a few milliseconds for all of them costs nothing): ZPQ_SHIM_WINDOW_MS
int window_ms() { static const int ms = [] { const char* e = getenv("ZPQ_SHIM_WINDOW_MS"); const int v = e ? atoi(e) : 100; return v < 0 ? 0 : (v > 1000 ? 1000 : v); }(); return ms; }

int max_wait_ms() { static const int ms = [] { const char* e = getenv("ZPQ_SHIM_MAX_WAIT_MS"); const int v = e ? atoi(e) : 30000; return v < 0 ? 0 : (v > 600000 ? 600000 : v); }(); return ms; }

struct Batcher {
  std::mutex mu; std::condition_variable cv_submit, cv_done;
  std::vector<Req*> queue;
  zpq_ctx* ctx = nullptr; int ctx_rc = 1;
  bool started = false;

  void run() {
    for (;;) {
      std::vector<Req*> batch;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_submit.wait(lk, [&] { return !queue.empty(); });
        // give concurrently running compress / decompress threads time to add their blocks (zpaq keeps `-threads`
        // blocks in flight: `zpaq add -threads 512` hands over 1 023 blocks as fast as the reader thread fills them).
        // A launch is bound by the latency of ONE block (about ten seconds for a 1 MiB block at level 5, whatever the
        // batch size), so a batch should be everything the caller will supply: keep waiting in steps of
        // ZPQ_SHIM_WINDOW_MS (100 ms: zpaq's reader thread fragments and hashes at ~100 MB/s, a block every ~10 ms) while
        // the queue is still growing, at most ZPQ_SHIM_MAX_WAIT_MS (30 s) in all.  Measured: profiles/r02q_cli_threads.log
        const auto t0 = std::chrono::steady_clock::now();
        size_t seen = queue.size();
        for (;;) {
          lk.unlock();
          std::this_thread::sleep_for(std::chrono::milliseconds(window_ms()));
          lk.lock();
          const bool grew = queue.size() != seen;
          seen = queue.size();
          const auto waited = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
          if (!grew || waited >= max_wait_ms()) break;
        }
        batch.swap(queue);
      }
      if (ctx_rc == 1) ctx_rc = zpq_create(0, &ctx);
      // one zpq_compress_blocks call per distinct (method, dosha1)
      std::vector<bool> used(batch.size(), false);
      // every decode request of the batch in one zpq_decode_blocks call
      {
        std::vector<size_t> idx;
        for (size_t j = 0; j < batch.size(); ++j) if (batch[j]->decode) { used[j] = true; idx.push_back(j); }
        if (!idx.empty()) {
          std::vector<zpq_job> jobs(idx.size());
          for (size_t k = 0; k < idx.size(); ++k) {
            Req* r = batch[idx[k]]; zpq_job& j = jobs[k];
            memset(&j, 0, sizeof j);
            j.hdr = r->hdr; j.hlen = r->hlen; j.nseg = (int)r->seg_len->size(); j.data = r->data; j.len = r->len;
            j.seg_len = r->seg_len->data(); r->seg_out.assign(r->seg_len->size(), 0);
            j.out = r->out.data(); j.out_cap = r->out.size(); j.seg_out = r->seg_out.data(); j.status = ZPQ_E_CUDA;
          }
          int rc = ctx_rc;
          if (ctx_rc == ZPQ_OK) rc = zpq_decode_blocks(ctx, (int)jobs.size(), jobs.data());
          for (size_t k = 0; k < idx.size(); ++k) {
            Req* r = batch[idx[k]];
            r->status = (rc != ZPQ_OK && jobs[k].status == ZPQ_OK) ? rc : jobs[k].status;
            r->out_len = jobs[k].out_len;
          }
        }
      }
      for (size_t i = 0; i < batch.size(); ++i) {
        if (used[i]) continue;
        std::vector<size_t> idx;
        for (size_t j = i; j < batch.size(); ++j)
          if (!used[j] && batch[j]->method == batch[i]->method && batch[j]->dosha1 == batch[i]->dosha1) { used[j] = true; idx.push_back(j); }
        const int n = (int)idx.size();
        std::vector<const uint8_t*> data(n); std::vector<uint64_t> len(n), cap(n), olen(n);
        std::vector<const char*> fn(n), cm(n); std::vector<uint8_t*> outp(n); std::vector<int32_t> st(n, ZPQ_E_CUDA);
        for (int k = 0; k < n; ++k) {
          Req* r = batch[idx[k]];
          data[k] = r->data; len[k] = r->len; fn[k] = r->filename; cm[k] = r->comment;
          r->out.resize(zpq_compress_bound(r->len) + 70000);
          outp[k] = r->out.data(); cap[k] = r->out.size();
        }
        int rc = ctx_rc;
        if (ctx_rc == ZPQ_OK)
          rc = zpq_compress_blocks(ctx, n, batch[i]->method.c_str(), data.data(), len.data(), fn.data(), cm.data(),
                                   batch[i]->dosha1 ? 1 : 0, outp.data(), cap.data(), olen.data(), st.data());
        for (int k = 0; k < n; ++k) {
          Req* r = batch[idx[k]];
          r->status = (rc != ZPQ_OK && st[k] == ZPQ_OK) ? rc : st[k];
          r->out_len = olen[k];
        }
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        for (Req* r : batch) r->done = true;
      }
      cv_done.notify_all();
    }
  }

  void enqueue(Req* r) {
    std::unique_lock<std::mutex> lk(mu);
    if (!started) { started = true; std::thread([this] { run(); }).detach(); }
    queue.push_back(r);
    cv_submit.notify_one();
  }
  void wait(Req* r) {
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [&] { return r->done; });
  }
  void submit(Req* r) { enqueue(r); wait(r); }
};

Batcher& batcher() { static Batcher* b = new Batcher(); return *b; }

// which methods the GPU library takes is decided by the library itself: zpq_model_from_method for the methods that feed the
// model with the payload, zpq_block_plan for those that put a transform in front of it (BWT, E8E9: zpq_bwt.cu).  Store and
// LZ77 selections come back as ZPQ_E_METHOD.
int probe_method(const char* method, const uint8_t* data, uint64_t n, uint8_t* hdr, int cap, int* hlen) {
  int rc = zpq_model_from_method(method, data, n, hdr, cap, hlen);
  if (rc == ZPQ_E_METHOD) {
    static thread_local uint8_t prefix[70000];
    int prep = 0, plen = 0;
    const int32_t g[2] = {0, 0};
    const int rp = zpq_block_plan(method, n, g, hdr, cap, hlen, &prep, prefix, (int)sizeof(prefix), &plen);
    // BWT (3, 7) and E8E9 in front of a context model (4) run on the device at batch speed.  The LZ77 parsers (1, 2, 5, 6) are
    // built too, but they are sequential per block (one GPU thread each): for zpaq's 16 MiB default blocks and its small
    // index blocks libzpaq's own parser is the better choice, so they stay there unless ZPQ_SHIM_LZ77=1 asks for the library
    static const bool lz_too = [] { const char* e = getenv("ZPQ_SHIM_LZ77"); return e && atoi(e) != 0; }();
    if (rp == ZPQ_OK && prep != 0 && (lz_too || (prep & 3) == 3 || prep == 4)) rc = ZPQ_OK;
  }
  return rc;
}

// does the CUDA library code blocks of this (valid, modelled) header?
bool gpu_takes(const uint8_t* hdr, int hlen) {
  uint32_t info[8];
  return zpq_model_describe(hdr, hlen, info) == ZPQ_OK;
}

}  // namespace

void compressBlock(StringBuffer* in, Writer* out, const char* method, const char* filename, const char* comment, bool dosha1) {
  uint8_t hdr[70000]; int hlen = 0;
  const int probe = probe_method(method, in->data(), in->size(), hdr, (int)sizeof(hdr), &hlen);
  if (probe == ZPQ_E_METHOD || probe == ZPQ_E_UNSUPPORTED || (probe == ZPQ_OK && hdr[6] == 0)) {
    SHIM_LOG("compressBlock %d bytes method %s -> reference CPU code (store / LZ77 selection)", (int)in->size(), method);
    compressBlock_cpu(in, out, method, filename, comment, dosha1);     // LZ77 / stored: not on the hot path, no transform kernel for them
    return;
  }
  if (probe == ZPQ_OK && !gpu_takes(hdr, hlen))       // a modelled block the kernels cannot take: say so, never code it on the CPU
    error("context model not supported by the B200 core");
  if (probe != ZPQ_OK) error(zpq_strerror(probe));
  Req r;
  r.data = in->data(); r.len = in->size(); r.method = method; r.filename = filename; r.comment = comment; r.dosha1 = dosha1;
  batcher().submit(&r);
  SHIM_LOG("compressBlock %d bytes method %s -> GPU, status %d, %d bytes", (int)in->size(), method, r.status, (int)r.out_len);
  if (r.status != ZPQ_OK) error(zpq_strerror(r.status));               // error() does not return (zpaq.cpp:123-126)
  out->write((const char*)r.out.data(), (int)r.out_len);
}


// ------------------------------------------------------------------------------------------------
// libzpaq::compress(Reader*, Writer*, method, ...) (libzpaq.cpp:3005-3027): the input is cut into blocks of
// (2^(20+N) - 4096) bytes, N = the digits after the level in `method`.  The reference codes one block at a time;
// here up to ZPQ_SHIM_BATCH blocks (default 256, at most ~2 GiB of input) are read ahead and coded as ONE GPU batch,
// then written in order.  Blocks of methods outside the context-model family go to the reference's own code.
// ------------------------------------------------------------------------------------------------
void compress(Reader* in, Writer* out, const char* method, const char* filename, const char* comment, bool dosha1) {
  // block size exponent: up to two decimal digits right after the level character, default 4, at most 11
  int bs = 4;
  const char* digits = (method && method[0]) ? method + 1 : "";
  if (isdigit((unsigned char)digits[0])) {
    bs = digits[0] - '0';
    if (isdigit((unsigned char)digits[1])) bs = 10 * bs + (digits[1] - '0');
  }
  bs = std::min(bs, 11);
  const size_t block = ((size_t)1 << (20 + bs)) - 4096;
  size_t max_blocks = 256;
  if (const char* e = getenv("ZPQ_SHIM_BATCH")) { const long v = atol(e); if (v >= 1 && v <= 65536) max_blocks = (size_t)v; }
  max_blocks = std::max<size_t>(1, std::min(max_blocks, ((size_t)2 << 30) / block));
  bool more = in != nullptr;
  while (more) {
    std::vector<StringBuffer*> bufs;
    while (more && bufs.size() < max_blocks) {
      StringBuffer* sb = new StringBuffer(block);
      sb->write(0, (int)block);
      const int n = in->read((char*)sb->data(), (int)block);
      if (n <= 0) { delete sb; more = false; break; }
      sb->resize(n);
      bufs.push_back(sb);                       // one block per read() call, short or not, exactly as the reference cuts them
    }
    std::vector<Req> reqs(bufs.size());
    std::vector<char> on_gpu(bufs.size(), 0);
    for (size_t i = 0; i < bufs.size(); ++i) {
      uint8_t hdr[70000]; int hlen = 0;
      const int probe = probe_method(method, bufs[i]->data(), bufs[i]->size(), hdr, (int)sizeof(hdr), &hlen);
      if (probe == ZPQ_OK && hdr[6] > 0 && !gpu_takes(hdr, hlen)) {
        for (size_t k = 0; k < i; ++k) if (on_gpu[k]) batcher().wait(&reqs[k]);
        for (StringBuffer* sb : bufs) delete sb;
        error("context model not supported by the B200 core");
      }
      if (probe == ZPQ_OK && hdr[6] > 0) {
        Req& r = reqs[i];
        r.data = bufs[i]->data(); r.len = bufs[i]->size(); r.method = method;
        r.filename = i == 0 ? filename : nullptr; r.comment = i == 0 ? comment : nullptr; r.dosha1 = dosha1;
        on_gpu[i] = 1;
        batcher().enqueue(&r);
      }
      else if (probe != ZPQ_OK && probe != ZPQ_E_METHOD && probe != ZPQ_E_UNSUPPORTED) {
        for (size_t k = 0; k < i; ++k) if (on_gpu[k]) batcher().wait(&reqs[k]);      // nothing may outlive this frame
        for (StringBuffer* sb : bufs) delete sb;
        error(zpq_strerror(probe));
      }
    }
    for (size_t i = 0; i < bufs.size(); ++i) if (on_gpu[i]) batcher().wait(&reqs[i]);
    SHIM_LOG("compress: %d blocks of up to %d bytes read ahead, method %s, %d of them -> GPU", (int)bufs.size(), (int)block, method,
             (int)std::count(on_gpu.begin(), on_gpu.end(), (char)1));
    int bad = ZPQ_OK;
    for (size_t i = 0; i < bufs.size() && bad == ZPQ_OK; ++i) {
      if (on_gpu[i]) {
        if (reqs[i].status != ZPQ_OK) bad = reqs[i].status;
        else out->write((const char*)reqs[i].out.data(), (int)reqs[i].out_len);
      }
      else compressBlock_cpu(bufs[i], out, method, i == 0 ? filename : nullptr, i == 0 ? comment : nullptr, dosha1);
    }
    for (StringBuffer* sb : bufs) delete sb;
    if (bad != ZPQ_OK) error(zpq_strerror(bad));
    filename = nullptr; comment = nullptr;
  }
}

// ------------------------------------------------------------------------------------------------
// Decompresser::decompress / readSegmentEnd (libzpaq.cpp:2315-2373)
// ------------------------------------------------------------------------------------------------
}  // namespace libzpaq
// the reference's own member functions under their objcopy-renamed symbols (`this` is the first argument)
extern "C" bool zpq_ref_decompress(libzpaq::Decompresser* self, int n);
extern "C" void zpq_ref_readSegmentEnd(libzpaq::Decompresser* self, char* sha1string);
namespace libzpaq {

namespace {

// What the shim remembers about a Decompresser between calls.  libzpaq.h is unchanged, so the record lives in a
// side table keyed by the object's address; an object is (re)started whenever decompress() sees FIRSTSEG.
struct DState {
  bool gpu = false;                    // this block is decoded by the GPU library
  std::vector<uint8_t> hdr;            // block header as stored (ZPAQL::write layout)
  std::vector<uint8_t> coded;          // coded segments of the block so far, each with its four zeros
  std::vector<uint64_t> seg_len;
  std::vector<uint8_t> plain;          // model output of the current segment (post-processor prefix included)
  uint64_t pos = 0;                    // bytes of `plain` already given to the post-processor
  bool have_plain = false;             // current segment decoded, its coded bytes consumed from the Reader
  int marker = -2;                     // byte that followed the segment's zeros (253 / 254), read ahead; -2 = none
};
std::mutex g_dmu;
std::map<const Decompresser*, DState*> g_dstate;

DState* dstate(const Decompresser* d) {
  std::lock_guard<std::mutex> lk(g_dmu);
  DState*& s = g_dstate[d];
  if (!s) s = new DState();
  return s;
}

}  // namespace

bool Decompresser::decompress(int n) {
  if (decode_state == SKIP) error("decompression after skipped segment");
  DState* S = dstate(this);
  if (decode_state == FIRSTSEG) {
    // new block: does the GPU library take its model?  header = COMP part + HCOMP part without the guard gap
    // between them (ZPAQL::write, libzpaq.cpp:866-884)
    *S = DState();
    if (z.header.isize() > 6 && z.cend > 0 && z.hbegin >= z.cend && z.hend >= z.hbegin) {
      for (int i = 0; i < z.cend; ++i) S->hdr.push_back(z.header[i]);
      for (int i = z.hbegin; i < z.hend; ++i) S->hdr.push_back(z.header[i]);
      uint32_t info[8];
      const int drc = z.header[6] > 0 ? zpq_model_describe(S->hdr.data(), (int)S->hdr.size(), info) : ZPQ_E_UNSUPPORTED;
      S->gpu = drc == ZPQ_OK;
      // a modelled block the kernels cannot take is an error, not a reason to decode it on the CPU (invalid headers
      // still go to the reference's Predictor::init, which raises its own message)
      if (z.header[6] > 0 && drc == ZPQ_E_UNSUPPORTED) error("context model not supported by the B200 core");
    }
    if (S->gpu) { pp.init(z.header[4], z.header[5]); decode_state = SEG; }     // dec.init() is the CPU model: not needed
    SHIM_LOG("Decompresser: block with %d components -> %s", z.header.isize() > 6 ? (int)z.header[6] : -1, S->gpu ? "GPU" : "reference CPU code");
  }
  if (!S->gpu) return zpq_ref_decompress(this, n);

  if (!S->have_plain) {
    // the segment's coded bytes: everything up to the end-of-segment marker.  The coder never emits four zeros in a
    // row inside the data (libzpaq.cpp:2414), but the last bytes of the range flush may be zero, so the run that ends a
    // segment is 4..7 zeros long: after four zeros keep reading zeros, the first other byte is the 253 / 254 marker
    // (Decoder::skip: "might be more than 4", libzpaq.cpp:2165).  The GPU decoder wants all of them (it stops four
    // bytes after the flush, at the end of its input).
    const uint64_t start = S->coded.size();
    int zeros = 0, c = 0;
    while (zeros < 4) {
      c = dec.get();
      if (c < 0) error("unexpected end of file");
      S->coded.push_back((uint8_t)c);
      zeros = c == 0 ? zeros + 1 : 0;
    }
    while ((c = dec.get()) == 0) S->coded.push_back(0);
    S->marker = c;                                      // consumed here, parsed by readSegmentEnd below
    S->seg_len.push_back(S->coded.size() - start);
    // decode every segment of the block read so far (model and coder state carry over between segments,
    // libzpaq.cpp:2889); the output size is not known in advance: grow the buffer until it fits
    Req r;
    r.decode = true; r.hdr = S->hdr.data(); r.hlen = (int)S->hdr.size();
    r.data = S->coded.data(); r.len = S->coded.size(); r.seg_len = &S->seg_len;
    uint64_t cap = S->coded.size() * 4 + (1u << 16);
    for (;;) {
      r.out.assign(cap, 0); r.done = false; r.status = 1;
      batcher().submit(&r);
      if (r.status != ZPQ_E_OUTPUT_FULL || cap >= ((uint64_t)1 << 32)) break;
      cap *= 4;
    }
    SHIM_LOG("Decompresser: segment %d of the block, %d coded bytes -> GPU, status %d, %d bytes", (int)S->seg_len.size(), (int)S->seg_len.back(), r.status, (int)r.out_len);
    if (r.status != ZPQ_OK) error(zpq_strerror(r.status));
    uint64_t before = 0;
    for (size_t k = 0; k + 1 < r.seg_out.size(); ++k) before += r.seg_out[k];
    S->plain.assign(r.out.begin() + before, r.out.begin() + before + r.seg_out.back());
    S->pos = 0; S->have_plain = true;
  }

  // the reference's loop, fed from the decoded segment: load PCOMP (if any) into the post-processor, then n bytes
  while ((pp.getState() & 3) != 1) {
    if (S->pos >= S->plain.size()) error("unexpected end of file");
    pp.write(S->plain[S->pos++]);
  }
  while (n) {
    if (S->pos >= S->plain.size()) {
      pp.write(-1);
      state = SEGEND;
      S->have_plain = false; S->plain.clear(); S->plain.shrink_to_fit();
      return false;
    }
    pp.write(S->plain[S->pos++]);
    if (n > 0) --n;
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// libzpaq::decompress(Reader*, Writer*) (libzpaq.cpp:2376-2389).  The whole archive is read and every block
// is decoded in ONE GPU batch (zpq_decompress_archive); an archive with blocks outside that call's scope
// (PCOMP post-processing, stored blocks) goes block by block through Decompresser above instead.
// ------------------------------------------------------------------------------------------------
namespace {
struct MemReader : public Reader {
  const uint8_t* p; size_t n, pos;
  MemReader(const uint8_t* p_, size_t n_) : p(p_), n(n_), pos(0) {}
  int get() { return pos < n ? p[pos++] : -1; }
  int read(char* buf, int k) { const size_t m = std::min<size_t>((size_t)std::max(k, 0), n - pos); memcpy(buf, p + pos, m); pos += m; return (int)m; }
};
zpq_ctx* stream_ctx() {          // null without a usable device
  static std::mutex mu; static zpq_ctx* c = nullptr; static int rc = 1;
  std::lock_guard<std::mutex> lk(mu);
  if (rc == 1) rc = zpq_create(0, &c);
  return rc == ZPQ_OK ? c : nullptr;
}
}  // namespace

void decompress(Reader* in, Writer* out) {
  std::vector<uint8_t> arc;
  for (;;) {
    const size_t at = arc.size();
    arc.resize(at + (1u << 20));
    const int n = in ? in->read((char*)arc.data() + at, 1 << 20) : 0;
    arc.resize(at + (n > 0 ? (size_t)n : 0));
    if (n <= 0) break;
  }
  if (arc.empty()) return;
  // Without a device the archive goes block by block through Decompresser: blocks that never needed the GPU
  // (stored, LZ77 index blocks) still decode, a context-model block raises "CUDA error" there -- no CPU fallback.
  if (zpq_ctx* ctx = stream_ctx()) {
    std::vector<uint8_t> plain(arc.size() * 4 + (1u << 20));
    uint64_t n = 0; int nb = 0;
    int rc = zpq_decompress_archive(ctx, arc.data(), arc.size(), plain.data(), plain.size(), &n, &nb, 0);
    if (rc == ZPQ_E_OUTPUT_FULL && n > plain.size()) {          // n = the size it needs
      plain.resize(n);
      rc = zpq_decompress_archive(ctx, arc.data(), arc.size(), plain.data(), plain.size(), &n, &nb, 0);
    }
    SHIM_LOG("decompress: archive of %d bytes, %d blocks -> GPU batch, status %d, %d bytes", (int)arc.size(), nb, rc, (int)n);
    if (rc == ZPQ_OK) {
      for (uint64_t at = 0; at < n; at += (1u << 30)) out->write((const char*)plain.data() + at, (int)std::min<uint64_t>(n - at, 1u << 30));
      return;
    }
    if (rc != ZPQ_E_UNSUPPORTED && rc != ZPQ_E_HEADER) error(zpq_strerror(rc));
  }
  MemReader mr(arc.data(), arc.size());
  Decompresser d;
  d.setInput(&mr);
  d.setOutput(out);
  while (d.findBlock())
    while (d.findFilename()) { d.readComment(); d.decompress(); d.readSegmentEnd(); }
}

void Decompresser::readSegmentEnd(char* sha1string) {
  DState* S = dstate(this);
  if (decode_state == FIRSTSEG) *S = DState();      // no decompress() in this block yet: whatever the table holds is an older object's
  if (S->gpu && S->marker != -2) {
    // the segment's coded bytes, its zeros and the marker byte are already consumed (decompress above).  A segment
    // left half way (state DATA) counts as skipped: as in the reference no further decompress() may follow in this block.
    if (state == DATA) decode_state = SKIP;
    S->have_plain = false; S->plain.clear();
    int c = S->marker;
    S->marker = -2;
    state = FILENAME;
    if (c == 254) { if (sha1string) sha1string[0] = 0; }
    else if (c == 253) {
      if (sha1string) sha1string[0] = 1;
      for (int i = 1; i <= 20; ++i) { c = dec.get(); if (sha1string) sha1string[i] = (char)c; }
    }
    else error("missing end of segment marker");
    return;
  }
  zpq_ref_readSegmentEnd(this, sha1string);
}

}  // namespace libzpaq
