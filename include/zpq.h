/* include/zpq.h -- C ABI of the B200-native context-mixing core (libzpq_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of libzpaq 7.15: the per-bit
 * Predictor::predict()/update() loop over the COMP chain, fused with the
 * Encoder/Decoder arithmetic coder and the ZPAQL HCOMP interpreter.  Each entry
 * point cites the reference interface it replaces (file:line in zpaq/zpaq).
 * Plain C types, caller-allocated buffers, int status codes, no exceptions, no
 * torch types.  INTEGRATION.md shows the libzpaq-side binding.
 *
 * Every compute entry point runs on the GPU.  There is no CPU fallback: without
 * a usable device they return ZPQ_E_CUDA.
 */
#ifndef ZPQ_H
#define ZPQ_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZPQ_ABI_VERSION 1

/* Status codes.  zpq_strerror() maps them to the reference's own message
 * strings where one exists (the libzpaq shim passes them to libzpaq::error,
 * libzpaq.h:857, which must not return). */
enum {
  ZPQ_OK = 0,
  ZPQ_E_HEADER = -1,        /* malformed block header (ZPAQL::read, libzpaq.cpp:887-930)        */
  ZPQ_E_COMPONENT = -2,     /* invalid COMP arguments (Predictor::init, libzpaq.cpp:1768-1850)  */
  ZPQ_E_VM = -3,            /* "ZPAQL execution error" (ZPAQL::err, libzpaq.cpp:1265)           */
  ZPQ_E_OUTPUT_FULL = -4,   /* caller's output buffer too small                                 */
  ZPQ_E_CORRUPT = -5,       /* "archive corrupted" (Decoder::decode, libzpaq.cpp:2108)          */
  ZPQ_E_EOF = -6,           /* "unexpected end of file" (Decoder::decode, libzpaq.cpp:2121)     */
  ZPQ_E_NOMEM = -7,         /* "Out of memory" (Array<T>::resize, libzpaq.h:921)                */
  ZPQ_E_UNSUPPORTED = -8,   /* valid in libzpaq but not built here              */
  ZPQ_E_CUDA = -9,          /* CUDA runtime failure; zpq_last_error() has the text              */
  ZPQ_E_ARG = -10,          /* bad argument                                                     */
  ZPQ_E_EOS = -11,          /* "decoding end of stream" (Decoder::decompress, libzpaq.cpp:2134) */
  ZPQ_E_CHECKSUM = -12,     /* SHA-1 of decoded segment differs from the stored one             */
  ZPQ_E_METHOD = -13        /* method string outside the context-model family built here        */
};

typedef struct zpq_ctx zpq_ctx;

/* One block to code.  `hdr` is the block header exactly as stored in the
 * archive after "zPQ" level type: hsize[2] hh hm ph pm n COMP.. 0 HCOMP.. 0
 * (ZPAQL::write, libzpaq.cpp:866-884).
 *
 * Encode: `data` are the bytes that pass THROUGH the model, i.e. the
 *   post-processor prefix (PASS byte 0, or 1 len_lo len_hi pcomp..,
 *   Compressor::postProcess libzpaq.cpp:2888-2917) followed by the payload.
 *   With nseg>1, seg_len[] splits `data` into segments; each ends with EOS and
 *   model + coder state carry over (libzpaq.cpp:2889).  Output: the
 *   arithmetic-coded bytes of each segment back to back (incl. the 4 EOS flush
 *   bytes, excl. the 00 00 00 00 terminator), lengths in seg_out[].
 * Decode: `data` are the coded segments back to back, EACH FOLLOWED BY THE FOUR
 *   ZERO BYTES of its terminator (Decoder::decompress reads them,
 *   libzpaq.cpp:2129-2136); seg_len[] counts each segment's bytes INCLUDING
 *   those four.  Output: decoded bytes (prefix included) back to back, lengths
 *   in seg_out[].
 * seg_len == NULL means one segment of `len` bytes.  `status` receives the
 * per-block result; a failed block does not fail the batch. */
typedef struct zpq_job {
  const uint8_t* hdr;
  int32_t hlen;
  int32_t nseg;
  const uint8_t* data;
  uint64_t len;
  const uint64_t* seg_len;
  uint8_t* out;
  uint64_t out_cap;
  uint64_t* seg_out;      /* nseg entries (or 1) */
  uint64_t out_len;       /* total bytes written to out */
  int32_t status;
  int32_t reserved;
} zpq_job;

/* ---- lifetime ------------------------------------------------------------ */
int zpq_abi_version(void);
/* Bind a context to CUDA device `device`, upload the model-independent tables
 * (Predictor::init, libzpaq.cpp:1731-1750; StateTable, libzpaq.cpp:726-855). */
int zpq_create(int device, zpq_ctx** ctx);
void zpq_destroy(zpq_ctx* ctx);
const char* zpq_strerror(int status);
const char* zpq_last_error(const zpq_ctx* ctx);
/* Give the context a caller-owned device workspace (e.g. a torch tensor) to
 * hold the per-block model arenas; NULL/0 returns to internal cudaMalloc. */
int zpq_set_workspace(zpq_ctx* ctx, void* device_ptr, uint64_t bytes);
/* Use a caller-owned CUDA stream (cudaStream_t as void*); NULL = own stream. */
int zpq_set_stream(zpq_ctx* ctx, void* cuda_stream);

/* ---- model --------------------------------------------------------------- */
/* Device bytes one block of this model needs (replaces ZPAQL::memory(),
 * libzpaq.cpp:986-1006, with the GPU arena layout); validates COMP exactly as
 * Predictor::init does. */
int zpq_model_bytes(const uint8_t* hdr, int hlen, uint64_t* bytes);
/* Row C1 of SURVEY.md 8(a): libzpaq::compressBlock's level-4/5 model selection
 * (libzpaq.cpp:7580-7691) + makeConfig (6887-7535) + Compiler (2449-2770) for
 * the context-model family (no LZ77/BWT/E8E9 pre-processing): method string
 * ("4", "5", "50", "x0,0ci1,1,..." ...) and the block payload -> header bytes.
 * The payload is needed because level 5 analyses it (libzpaq.cpp:7655-7688). */
int zpq_model_from_method(const char* method, const uint8_t* data, uint64_t n,
                          uint8_t* hdr, int cap, int* hlen);
/* Diagnostics of the kernel plan for a header: info[0]=components, [1]=predict steps,
 * [2]=shared-memory bytes per block, [3]=1 if HCOMP runs as lane-parallel recipes
 * (0: bytecode interpreter), [4]=recipe rounds, [5]=1 if the fast kernel takes this model,
 * [6]=generic-kernel steps, [7]=HCOMP bytes. */
int zpq_model_describe(const uint8_t* hdr, int hlen, uint32_t info[8]);

/* ---- the hot path -------------------------------------------------------- */
/* Host buffers in, host buffers out (copies inside).  Replaces the loop
 * Compressor::compress -> Encoder::compress (libzpaq.cpp:2920-2944, 2419-2434)
 * for a whole batch of independent blocks. */
int zpq_encode_blocks(zpq_ctx* ctx, int nblk, zpq_job* jobs);
/* Replaces Decompresser::decompress -> Decoder::decompress
 * (libzpaq.cpp:2315-2343, 2127-2155). */
int zpq_decode_blocks(zpq_ctx* ctx, int nblk, zpq_job* jobs);
/* Same with job.data / job.out being DEVICE pointers already resident in HBM
 * (hdr, seg_len, seg_out stay host pointers).  Asynchronous work is finished
 * before return. */
int zpq_encode_blocks_dev(zpq_ctx* ctx, int nblk, zpq_job* jobs);
int zpq_decode_blocks_dev(zpq_ctx* ctx, int nblk, zpq_job* jobs);
/* Device time (ms, CUDA events on the context's stream) of the last batch:
 * [0] arena initialisation kernel, [1] coding kernel, [2] whole call. */
int zpq_last_timing(const zpq_ctx* ctx, float ms[3]);
/* Device time (ms) of the stages in front of the coder in the last zpq_compress_blocks[_dev] call (both are part of
 * zpq_last_timing's [0]): [0] SHA-1 + level-5 statistics, [1] E8E9 / BWT pre-processing. */
int zpq_last_prep_timing(const zpq_ctx* ctx, float ms[2]);
/* Algorithmic HBM bytes of the suffix-sorting kernels of that call (sort passes x 32 bytes per still-unsorted suffix + the
 * per-round bookkeeping), the numerator of their roofline. */
uint64_t zpq_last_transform_bytes(const zpq_ctx* ctx);
/* Number of kernel launches issued by this context so far. */
uint64_t zpq_launch_count(const zpq_ctx* ctx);
/* Diagnostics: blocks of the last batch coded by each kernel tier: [0] wavefront encoder (zpq_kernels5.cu),
 * [1] critical-first two-warp decoder (zpq_kernels7.cu), [2] round-1 two-warp decoder (zpq_kernels6.cu),
 * [3] one-warp fast kernel (zpq_kernels2.cu), [4] generic kernel (zpq_kernels.cu), [5] wide generic kernel (n > 32),
 * [6] lock-step decoder with lanes = blocks (zpq_kernels8.cu). */
int zpq_last_tiers(const zpq_ctx* ctx, uint32_t tiers[7]);

/* ---- whole blocks (framing kept from the reference, host side) ------------ */
/* Batched libzpaq::compressBlock (libzpaq.cpp:7543-7731): for each payload
 * writes one complete .zpaq block -- locator tag, "zPQ" header, one segment
 * (filename, comment = decimal size [+ " " + comment]), coded data, SHA-1
 * trailer, end-of-block -- bit-identical to the reference for every method
 * compressBlock takes: digit methods "L[B][,R,T]" of levels 0-5 (store, LZ77,
 * BWT, context mixing, each with E8E9 for exe block types: libzpaq.cpp:7596-7691)
 * and "x..." methods (makeConfig, 6887-7535), with the pre-processing of
 * LZBuffer / e8e9() (6450-6883) done on the device.  Not taken: LZ77 with a
 * second context order ("x..." with $4 > 0) -> ZPQ_E_METHOD.  out[i] must hold
 * out_cap[i] bytes; out_len[i] returns the block size.  filenames/comments may
 * be NULL. */
int zpq_compress_blocks(zpq_ctx* ctx, int nblk, const char* method,
                        const uint8_t* const* data, const uint64_t* len,
                        const char* const* filenames, const char* const* comments, int dosha1,
                        uint8_t* const* out, const uint64_t* out_cap, uint64_t* out_len,
                        int32_t* status);
/* The same with payloads and output blocks in HBM (block sharding: raw blocks arrive over NCCL, coded blocks leave the
 * same way).  data[i] / out[i] are device pointers; every other argument is host memory.  SHA-1 (libzpaq.cpp:101-177) and
 * the level-5 payload statistics (libzpaq.cpp:7655-7688) are computed on the device as well, so the host never reads the
 * payload bytes. */
int zpq_compress_blocks_dev(zpq_ctx* ctx, int nblk, const char* method,
                            const uint8_t* const* d_data, const uint64_t* len,
                            const char* const* filenames, const char* const* comments, int dosha1,
                            uint8_t* const* d_out, const uint64_t* out_cap, uint64_t* out_len,
                            int32_t* status);
/* zpq_model_from_method with the payload statistics already taken (zpq_analyze_dev): gaps[2] as returned there. */
int zpq_header_from_stats(const char* method, uint64_t n, const int32_t gaps[2], uint8_t* hdr, int cap, int* hlen);
/* What zpq_compress_blocks does with a method, for callers that want to see or reuse the plan: the header bytes, the
 * pre-processing of the payload (makeConfig's second number, libzpaq.cpp:6894: 0 none, 3 BWT, 4 E8E9, 7 E8E9 then BWT --
 * LZBuffer level 3 and e8e9(), libzpaq.cpp:6450-6459, 6649-6658) and the bytes that pass through the model in front of
 * the (transformed) payload: "0" (PASS) or "1, len, PCOMP program" (Compressor::postProcess, libzpaq.cpp:2888-2915).
 * prep may also be 1 / 2 (LZ77 with variable-length / byte codes) and 5 / 6 (the same behind E8E9); a header with n = 0
 * components means the bytes are stored (Encoder::compress without a model, libzpaq.cpp:2436-2445). */
int zpq_block_plan(const char* method, uint64_t n, const int32_t gaps[2], uint8_t* hdr, int hdr_cap, int* hlen,
                   int* prep, uint8_t* prefix, int prefix_cap, int* prefix_len);
/* Device-side analysis of payloads in HBM: sha1 = 20 bytes per block (or NULL), gaps = 2 ints per block (the dominant
 * byte-gap periods level 5 models, 0 = none; or NULL).  Results are host arrays. */
int zpq_analyze_dev(zpq_ctx* ctx, int nblk, const uint8_t* const* d_data, const uint64_t* len,
                    uint8_t* sha1, int32_t* gaps);
/* Framing scan of an in-memory archive (Decompresser::findBlock / findFilename / readComment / readSegmentEnd and
 * Decoder::skip, libzpaq.cpp:2158-2181, 2247-2373): one record per segment, for callers that shard blocks themselves
 * and decode with zpq_decode_blocks[_dev].  No device is involved.  coded_len includes the terminating zeros. */
typedef struct zpq_segment {
  uint32_t block, segment;      /* block index in the archive, segment index in the block */
  uint64_t hdr_off;             /* block header (hsize[2] ... HCOMP END) */
  uint32_t hlen, components;    /* header bytes; number of components (0 = stored block) */
  uint64_t coded_off, coded_len;
  int64_t size_hint;            /* decimal prefix of the segment comment (what compressBlock writes), -1 if none */
  int32_t has_sha1; uint8_t sha1[20];
} zpq_segment;
int zpq_scan_archive(const uint8_t* arc, uint64_t arc_len, zpq_segment* segs, int cap, int* nsegs, int* nblocks);
/* Batched libzpaq::decompress (libzpaq.cpp:2376-2389) over one in-memory
 * archive: finds every block (Decompresser::findBlock, libzpaq.cpp:2247),
 * decodes all modelled blocks concurrently on the GPU, post-processes on the
 * host (PASS prefix stripped; a PCOMP program -- BWT / LZ77 / E8E9 inverse
 * transforms -- is run by the host ZPAQL machine; unmodelled level-2 blocks are
 * unpacked on the host), verifies stored SHA-1s and writes the payloads back to
 * back in archive order. */
int zpq_decompress_archive(zpq_ctx* ctx, const uint8_t* arc, uint64_t arc_len,
                           uint8_t* out, uint64_t out_cap, uint64_t* out_len, int* nblocks,
                           int verify_sha1);
/* Host-side PostProcessor (libzpaq.cpp:2183-2241) for callers that bind at the hot-path level: `model_out` is what
 * zpq_decode_blocks returned for one block (nseg segments, lengths seg_len[]).  Its first byte is 0 (PASS: the rest is
 * the payload) or 1 (PROG: a PCOMP program follows and every later byte is run through it).  Writes the payload of all
 * segments back to back (out_len = total, even when it exceeds out_cap -> ZPQ_E_OUTPUT_FULL); seg_out[] (optional)
 * receives the per-segment payload sizes.  hdr = the block header (for its ph / pm sizes).  No device is involved. */
int zpq_postprocess_block(const uint8_t* hdr, int hlen, const uint8_t* model_out, const uint64_t* seg_len, int nseg,
                          uint8_t* out, uint64_t out_cap, uint64_t* out_len, uint64_t* seg_out);
/* Upper bound of a compressed block for `len` payload bytes. */
uint64_t zpq_compress_bound(uint64_t len);

#ifdef __cplusplus
}
#endif
#endif /* ZPQ_H */
