// libzpaq_shim.cpp -- link-compatible replacement of libzpaq::compressBlock for the UNCHANGED zpaq.cpp.
//
// Compiled against the unchanged reference header (libzpaq.h, never copied into this repo; the build
// recipe oracle/Makefile adds -I/root/reference) and linked in front of the reference library object,
// whose own compressBlock is renamed to libzpaq::compressBlock_cpu by objcopy.  zpaq.cpp's compress
// threads (zpaq.cpp:1918-1965) call this function, one block at a time; concurrent calls are coalesced
// into one GPU batch by a small submitter so that the device sees many blocks per launch.
//
//   * context-model methods (-method 4, 5, x0,0c...)  -> zpq_compress_blocks (CUDA, bit-identical bytes)
//   * methods that need LZ77/BWT/E8E9 pre-processing   -> the reference's own CPU code (out of scope of
//     the hot path)
//   * a context-model method without a usable GPU      -> libzpaq::error("CUDA error"): no CPU fallback
//     of the hot path exists in libzpq_b200.so.
#include "libzpaq.h"
#include "../../include/zpq.h"
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <chrono>

namespace libzpaq {

// the reference implementation under its objcopy-renamed symbol
void compressBlock_cpu(StringBuffer* in, Writer* out, const char* method, const char* filename, const char* comment, bool dosha1);

namespace {

struct Req {
  const uint8_t* data; uint64_t len; std::string method; const char* filename; const char* comment; bool dosha1;
  std::vector<uint8_t> out; uint64_t out_len = 0; int status = 1; bool done = false;
};

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
        // give concurrently running compress threads a moment to add their blocks (zpaq keeps
        // `threads` blocks in flight), then take everything that is queued
        lk.unlock();
        std::this_thread::sleep_for(std::chrono::milliseconds(2));
        lk.lock();
        batch.swap(queue);
      }
      if (ctx_rc == 1) ctx_rc = zpq_create(0, &ctx);
      // one zpq_compress_blocks call per distinct (method, dosha1)
      std::vector<bool> used(batch.size(), false);
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

  void submit(Req* r) {
    std::unique_lock<std::mutex> lk(mu);
    if (!started) { started = true; std::thread([this] { run(); }).detach(); }
    queue.push_back(r);
    cv_submit.notify_one();
    cv_done.wait(lk, [&] { return r->done; });
  }
};

Batcher& batcher() { static Batcher* b = new Batcher(); return *b; }

}  // namespace

void compressBlock(StringBuffer* in, Writer* out, const char* method, const char* filename, const char* comment, bool dosha1) {
  // which methods the GPU core takes is decided by the library itself (zpq_model_from_method)
  uint8_t hdr[70000]; int hlen = 0;
  const int probe = zpq_model_from_method(method, in->data(), in->size(), hdr, (int)sizeof(hdr), &hlen);
  if (probe == ZPQ_E_METHOD || probe == ZPQ_E_UNSUPPORTED || (probe == ZPQ_OK && (hdr[6] == 0 || hdr[6] > 32))) {
    compressBlock_cpu(in, out, method, filename, comment, dosha1);     // LZ77 / BWT / stored: not on the hot path
    return;
  }
  if (probe != ZPQ_OK) error(zpq_strerror(probe));
  Req r;
  r.data = in->data(); r.len = in->size(); r.method = method; r.filename = filename; r.comment = comment; r.dosha1 = dosha1;
  batcher().submit(&r);
  if (r.status != ZPQ_OK) error(zpq_strerror(r.status));               // error() does not return (zpaq.cpp:123-126)
  out->write((const char*)r.out.data(), (int)r.out_len);
}

}  // namespace libzpaq
