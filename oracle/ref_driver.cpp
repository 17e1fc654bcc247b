// oracle/ref_driver.cpp -- C-callable wrapper around the UNMODIFIED reference
// libzpaq (TEST INFRASTRUCTURE; never linked into the product library).
//
// Built by oracle/Makefile together with /root/reference/libzpaq.cpp (compiled
// in place, outputs only under oracle/_ref/).  Used to
//   * generate the golden vectors in tests/golden/ (oracle/make_golden.py),
//   * validate the C restatement oracle/zpq_oracle.c bit for bit,
//   * time the reference on the host cores (bench.py --impl reference and the
//     cpu_baseline leg).
// Reference API used: libzpaq::compressBlock (libzpaq.h:1500),
// libzpaq::Decompresser (libzpaq.h:1243-1264), libzpaq::Compressor
// (libzpaq.h:1340-1371), libzpaq::makeConfig (libzpaq.cpp:6887, external
// linkage).  `#define private public` is used only to step
// Predictor::predict()/update() for per-bit traces.
#define private public
#include "libzpaq.h"
#undef private
#include <stdexcept>
#include <string>
#include <vector>
#include <thread>
#include <atomic>
#include <chrono>
#include <cstring>
#include "gen.h"

namespace libzpaq {
std::string makeConfig(const char* method, int args[]);
// The application supplies error(); it must not return (zpaq.cpp:123-126).
void error(const char* msg) { throw std::runtime_error(msg); }
}  // namespace libzpaq

namespace {
double* g_secs2 = nullptr;   // ref_bench_rt: where ref_bench leaves both phase times
std::string g_err;  // last error (callers are single threaded; bench workers only write on failure)

struct MemReader : public libzpaq::Reader {
  const unsigned char* p; size_t n, pos;
  MemReader(const void* d, size_t len) : p((const unsigned char*)d), n(len), pos(0) {}
  int get() { return pos < n ? p[pos++] : -1; }
  int read(char* buf, int len) {
    size_t k = n - pos; if (k > (size_t)len) k = len;
    memcpy(buf, p + pos, k); pos += k; return (int)k;
  }
};

struct VecWriter : public libzpaq::Writer {
  std::vector<unsigned char> v;
  void put(int c) { v.push_back((unsigned char)c); }
  void write(const char* buf, int n) { v.insert(v.end(), buf, buf + n); }
};

long long emit(const std::vector<unsigned char>& v, unsigned char* out, long long cap) {
  if ((long long)v.size() > cap) { g_err = "output buffer too small"; return -2; }
  if (!v.empty()) memcpy(out, v.data(), v.size());
  return (long long)v.size();
}
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// libzpaq::compressBlock on one in-memory block. Returns bytes written or <0.
long long ref_compress_block(const unsigned char* data, long long n, const char* method,
                             const char* filename, const char* comment, int dosha1,
                             unsigned char* out, long long cap) {
  try {
    libzpaq::StringBuffer in(n > 0 ? n : 1);
    if (n > 0) in.write((const char*)data, (int)n);
    VecWriter w;
    libzpaq::compressBlock(&in, &w, method, filename, comment, dosha1 != 0);
    return emit(w.v, out, cap);
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// Compress one block with an explicit ZPAQL config text (level==0) or a
// built-in model level 1..3 (min/mid/max.cfg, libzpaq.cpp:2793-2831).
// nseg>1 splits the data into nseg segments of one block (state persists).
long long ref_compress_config(const unsigned char* data, long long n, const char* config,
                              const int* args9, int level, int nseg, int dosha1,
                              unsigned char* out, long long cap) {
  try {
    VecWriter w;
    libzpaq::Compressor co;
    co.setOutput(&w);
    co.writeTag();
    int args[9] = {0};
    if (args9) memcpy(args, args9, sizeof(args));
    if (level > 0) co.startBlock(level);
    else co.startBlock(config, args);
    if (nseg < 1) nseg = 1;
    long long pos = 0;
    for (int s = 0; s < nseg; ++s) {
      long long len = (s == nseg - 1) ? n - pos : n / nseg;
      MemReader r(data + pos, (size_t)len);
      libzpaq::SHA1 sha1;
      sha1.write((const char*)data + pos, len);
      std::string fn = "seg" + std::to_string(s);
      co.startSegment(fn.c_str(), "");
      co.setInput(&r);
      co.compress();
      co.endSegment(dosha1 ? sha1.result() : 0);
      pos += len;
    }
    co.endBlock();
    return emit(w.v, out, cap);
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// libzpaq::decompress of a whole in-memory archive (all blocks, all segments).
long long ref_decompress(const unsigned char* arc, long long n, unsigned char* out, long long cap) {
  try {
    MemReader r(arc, (size_t)n);
    VecWriter w;
    libzpaq::decompress(&r, &w);
    return emit(w.v, out, cap);
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// makeConfig(method) -> ZPAQL source text and args[9].
long long ref_make_config(const char* method, int* args9, char* out, long long cap) {
  try {
    int args[9] = {0};
    std::string c = libzpaq::makeConfig(method, args);
    if (args9) memcpy(args9, args, sizeof(args));
    if ((long long)c.size() + 1 > cap) { g_err = "buffer too small"; return -2; }
    memcpy(out, c.c_str(), c.size() + 1);
    return (long long)c.size();
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// Per-bit trace: feed `data` (already including the PASS byte if wanted)
// through Predictor built from block header bytes `hdr` (hsize.. as stored in
// the archive after "zPQ" level type).  probs[k] = predict() before bit k.
long long ref_trace(const unsigned char* hdr, long long hlen, const unsigned char* data,
                    long long n, unsigned short* probs) {
  try {
    MemReader r(hdr, (size_t)hlen);
    libzpaq::ZPAQL z;
    z.read(&r);
    libzpaq::Predictor pr(z);
    pr.init();
    long long k = 0;
    for (long long i = 0; i < n; ++i)
      for (int b = 7; b >= 0; --b) {
        probs[k++] = (unsigned short)pr.predict();
        pr.update((data[i] >> b) & 1);
      }
    return k;
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// Multi-threaded timing of the reference on synthetic blocks, one block per
// thread at a time exactly as zpaq.cpp:1918-1965 / 2796-3068 do.
// mode 0 = compress, 1 = decompress (archives prepared untimed).
// Returns seconds of wall clock for the timed phase, <0 on error.
// total_out receives the sum of output sizes (archive bytes or payload bytes).
double ref_bench(const char* method, int nblocks, long long blocksize, int kind,
                 unsigned seed, int threads, int mode, long long* total_out) {
  try {
    if (threads < 1) threads = 1;
    std::vector<std::vector<unsigned char>> blocks(nblocks), arcs(nblocks);
    for (int i = 0; i < nblocks; ++i) {
      blocks[i].resize(blocksize);
      zpq_gen_block(kind, seed + i, blocks[i].data(), (size_t)blocksize);
    }
    std::atomic<int> next(0);
    std::atomic<long long> total(0);
    std::atomic<int> failed(0);
    auto comp = [&]() {
      try {
        for (int i; (i = next++) < nblocks;) {
          libzpaq::StringBuffer in(blocksize);
          in.write((const char*)blocks[i].data(), (int)blocksize);
          VecWriter w;
          libzpaq::compressBlock(&in, &w, method, "", 0, true);
          total += (long long)w.v.size();
          arcs[i].swap(w.v);
        }
      } catch (std::exception& e) { g_err = e.what(); failed = 1; }
    };
    auto decomp = [&]() {
      try {
        for (int i; (i = next++) < nblocks;) {
          MemReader r(arcs[i].data(), arcs[i].size());
          VecWriter w;
          libzpaq::decompress(&r, &w);
          if (w.v != blocks[i]) { failed = 2; return; }
          total += (long long)w.v.size();
        }
      } catch (std::exception& e) { g_err = e.what(); failed = 1; }
    };
    auto run = [&](bool dec) {
      next = 0; total = 0;
      std::vector<std::thread> th;
      auto t0 = std::chrono::steady_clock::now();
      for (int t = 0; t < threads; ++t) { if (dec) th.emplace_back(decomp); else th.emplace_back(comp); }
      for (auto& t : th) t.join();
      return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    double secs = run(false);
    if (g_secs2) { g_secs2[0] = secs; g_secs2[1] = 0; }
    if ((mode == 1 || mode == 2) && !failed) { secs = run(true); if (g_secs2) g_secs2[1] = secs; }
    if (failed) { if (failed == 2) g_err = "round trip mismatch"; return -1.0; }
    if (total_out) *total_out = total;
    return secs;
  } catch (std::exception& e) { g_err = e.what(); return -1.0; }
}

// One pass that times both directions: secs2[0] = compress, secs2[1] = decompress of the archives just written
// (bench.py's bounded CPU sample: every block is compressed once, not twice).
int ref_bench_rt(const char* method, int nblocks, long long blocksize, int kind, unsigned seed, int threads, double* secs2) {
  g_secs2 = secs2;
  const double s = ref_bench(method, nblocks, blocksize, kind, seed, threads, 2, nullptr);
  g_secs2 = nullptr;
  return s < 0 ? -1 : 0;
}

// Static tables of the reference after Predictor::init() (libzpaq.cpp:1731-1750)
// and the StateTable (libzpaq.cpp:726-855), for pinning the regenerated ones.
int ref_tables(unsigned char* ns1024, int* dt2k256, int* dt1024, unsigned short* squash4096,
               short* stretch32768) {
  try {
    // hsize=10: hh hm ph pm n=1 | const 128 | 0 | halt | 0
    static const unsigned char h2[] = {10, 0, 0, 0, 0, 0, 1, 1, 128, 0, 56, 0};
    MemReader r(h2, sizeof(h2));
    libzpaq::ZPAQL z;
    z.read(&r);
    libzpaq::Predictor pr(z);
    pr.init();
    memcpy(ns1024, pr.st.ns, 1024);
    memcpy(dt2k256, pr.dt2k, sizeof(pr.dt2k));
    memcpy(dt1024, pr.dt, sizeof(pr.dt));
    memcpy(squash4096, pr.squasht, sizeof(pr.squasht));
    memcpy(stretch32768, pr.stretcht, sizeof(pr.stretcht));
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

void ref_gen_block(int kind, unsigned seed, unsigned char* out, long long n) {
  zpq_gen_block(kind, seed, out, (size_t)n);
}

}  // extern "C"
