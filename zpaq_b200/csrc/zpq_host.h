// zpq_host.h -- internal host-side declarations shared by the C-ABI layer, the
// model compiler, the method->header generator and the framing code.
#ifndef ZPQ_HOST_H
#define ZPQ_HOST_H
#include <stdint.h>
#include <string>
#include <vector>
#include <algorithm>
#include "../../include/zpq.h"
#include "zpq_dev.h"

namespace zpq {

// zpq_model.cpp
void build_state_table(uint8_t ns[1024]);
void build_tables(ZTables* T);
int compile_model(const uint8_t* hdr, int hlen, ZModel* M, std::vector<uint8_t>* image, std::string* err);

// zpq_hcomp.cpp : HCOMP bytecode -> lane-parallel recipes (self-checked)
void build_recipes(ZModel* M, const uint8_t* prog, int prog_len);
void build_wave_plan(ZModel* M);   // after the pipe plan: flattened recipes for the wavefront encoder
void build_hash_plan(ZModel* M);
// PostProcessor (libzpaq.cpp:2183-2241) over a decoded block: PASS prefix stripped or PCOMP program run, per segment
int postprocess_block(int ph, int pm, const uint8_t* model_out, const uint64_t* seg_len, int nseg,
                      std::vector<uint8_t>* payload, std::vector<uint64_t>* seg_payload, std::string* err);   // after the wave plan: slot-scheduled hashes for the two-warp decoder

// zpq_method.cpp : method string (+ payload statistics) -> block header bytes
int header_from_method(const char* method, const uint8_t* data, uint64_t n,
                       std::vector<uint8_t>* hdr, std::string* err);
// the same with the payload statistics already taken (on the device: zpq_prep.cu): gaps[2] = dominant gaps, 0 = none
bool method_needs_gaps(const char* method);
void payload_gaps(const uint8_t* data, uint64_t n, int gaps[2]);
int header_from_stats(const char* method, uint64_t n, const int gaps[2], std::vector<uint8_t>* hdr, std::string* err);
// header + the pre-processing the payload needs (0 none, 3 BWT, 4 E8E9, 7 E8E9 then BWT) + the block size class
// lz[5]: the LZ77 numbers of the method (minimum match, second order, log2 bucket, table bits, look-ahead) when prep is 1, 2, 5 or 6
int block_plan(const char* method, uint64_t n, const int gaps[2], std::vector<uint8_t>* hdr, int* prep, int* size_class, int lz[5], std::string* err);

// zpq_pcomp.cpp : the post-processor programs of transformed blocks (bytes identical to makeConfig + Compiler)
int pcomp_program(int kind, int size_class, int min_match, std::vector<uint8_t>* code);
void pcomp_prefix(const std::vector<uint8_t>& code, std::vector<uint8_t>* prefix);

// zpq_frame.cpp : SHA-1 and block framing
struct Sha1 {
  uint32_t h[5]; uint64_t len; uint8_t buf[64]; int fill;
  Sha1() { reset(); }
  void reset();
  void update(const uint8_t* p, uint64_t n);
  void finish(uint8_t out[20]);
};
void sha1_of(const uint8_t* p, uint64_t n, uint8_t out[20]);

struct SegmentView {            // one segment located inside an archive
  uint64_t coded_off, coded_len;   // coded bytes INCLUDING the 4-zero terminator
  uint64_t comment_off, comment_len;
  bool has_sha1; uint8_t sha1[20];
};
struct BlockView {
  uint64_t hdr_off; int hlen;      // header bytes hsize[2]..HCOMP END
  int level;
  std::vector<SegmentView> segs;
  uint64_t end;                    // offset one past the 255 end-of-block byte
};
// Decompresser::findBlock/findFilename/readComment/readSegmentEnd (libzpaq.cpp:2247-2373) over memory
int scan_archive(const uint8_t* arc, uint64_t n, std::vector<BlockView>* blocks, std::string* err);
// Compressor::writeTag/startBlock/startSegment (libzpaq.cpp:2776-2883): bytes before the coded data
void write_block_prologue(std::vector<uint8_t>* out, const uint8_t* hdr, int hlen,
                          const char* filename, const char* comment);

}  // namespace zpq
#endif
