// zpq_dev.h -- structures shared by the host model compiler (zpq_model.cpp), the
// C-ABI layer (zpq_abi.cu) and the kernels (zpq_kernels.cu).
//
// HBM layout
//   workspace = [ block 0 arena | block 1 arena | ... ]     (arena_off, 256-byte aligned)
//   one arena = every table of one block's model, each table 128-byte aligned:
//       per component: cm (u32) / a16 (u16) / ht (u8) tables sized as
//       Predictor::init allocates them (libzpaq.cpp:1768-1850),
//       then the HCOMP machine: M (2^hm bytes), H (2^hh u32), R (256 u32)
//       (ZPAQL::init, libzpaq.cpp:1009-1024).
//   A block's tables are touched by exactly one warp, so they need no
//   coherence beyond program order.
#ifndef ZPQ_DEV_H
#define ZPQ_DEV_H
#include <stdint.h>

enum { ZT_NONE = 0, ZT_CONS, ZT_CM, ZT_ICM, ZT_MATCH, ZT_AVG, ZT_MIX2, ZT_MIX, ZT_ISSE, ZT_SSE };

#define ZPQ_MAX_LANES 32      // components per model on the warp-per-block path
#define ZPQ_MAX_STEPS 32
#define ZPQ_GUARD_LO 128      // zero guard before the HCOMP program (libzpaq.cpp:913)
#define ZPQ_GUARD_HI 132      // zero guard after it (operand fetches past a +128 jump)
#define ZPQ_STEP_PAR 0xff
// v2 step kinds (ZStep.kind): one uniform branch per step
enum { ZK_ISSE = 1, ZK_MIX, ZK_MIX2, ZK_SSE, ZK_AVG };

// One component as the kernel sees it (32 bytes).
struct ZComp {
  uint8_t type, a1, a2, a3, a4, a5;  // type byte and arguments as in the header
  uint8_t step;                      // index+1 of the step that produces p[i]; 0 = phase A
  uint8_t sizebits;                  // a1 copy (kept for clarity in find())
  uint32_t cm_off16;                 // cm / a16 table offset in the arena, 16-byte units
  uint32_t ht_off16;                 // ht table offset, 16-byte units
  uint32_t cm_mask;                  // entries-1 of the table indexed through cm(i) / size-1 for MIX, MIX2
  uint32_t ht_mask;                  // MATCH: buf bytes-1; ICM/ISSE: ht bytes-16
  uint32_t limit;                    // CM/SSE count limit (arg*4)
  uint32_t pad;
};

// One step of the dependent part of predict().
struct ZStep {
  uint8_t kind;     // ZPQ_STEP_PAR, or the index of the MIX component computed cooperatively
  uint8_t m;        // MIX: number of inputs
  uint8_t j0;       // MIX: first input
  uint8_t rate;     // MIX: learning rate
  uint32_t cm_off16;
};

// Fast-kernel plan: everything the per-bit loop needs sits in registers.
// Step word: kind | owner<<3 | j<<8 | k<<13 | aux<<19 | q<<27
//   ZK_ISSE : all ISSE lanes whose level equals (step index + 1) compute
//   ZK_MIX  : owner, j = first input, k = number of inputs, aux = rate, q = weight-register slot
//   ZK_MIX2 : owner, inputs j k        ZK_SSE : owner, input j, q = row-register slot
//   ZK_AVG  : owner, inputs j k, aux = weight
#define ZPQ_FAST_STEPS 16
#define ZPQ_FAST_MIX 3
#define ZPQ_FAST_SSE 2
#define ZPQ_FAST_RCP 12
#define ZSTEP_PACK(kind, owner, j, k, aux, q) \
  ((uint32_t)(kind) | (uint32_t)(owner) << 3 | (uint32_t)(j) << 8 | (uint32_t)(k) << 13 | (uint32_t)(aux) << 19 | (uint32_t)(q) << 27)
struct ZFast {
  uint32_t ok;                        // 1: the model fits the fast kernel; 0: generic kernel only
  uint32_t nsteps, nmix, nsse;
  uint32_t sd[ZPQ_FAST_STEPS];
  uint32_t mix_off16[ZPQ_FAST_MIX], mix_owner[ZPQ_FAST_MIX], mix_m[ZPQ_FAST_MIX], mix_j0[ZPQ_FAST_MIX], mix_rate[ZPQ_FAST_MIX];
  uint32_t sse_off16[ZPQ_FAST_SSE], sse_owner[ZPQ_FAST_SSE], sse_mask[ZPQ_FAST_SSE];
  uint32_t pw_bytes;                  // shared memory per block: all ICM (1 KB) and ISSE (2 KB) tables
  uint32_t smem_off[ZPQ_MAX_LANES];   // byte offset of the lane's table inside the block's shared region
  uint8_t level[ZPQ_MAX_LANES];       // ISSE lanes: dependency level (1 = fed by a phase-A component)
  uint32_t pair_ok;                   // two-warp decoder (zpq_kernels6.cu) usable: set after the wave plan
  uint32_t nlev;                      // number of ISSE levels
};


// Arena initialisation op (zpq_init_kernel).
enum { ZF_ZERO = 0, ZF_CONST32, ZF_SSE, ZF_ICM, ZF_ISSE, ZF_BYTE1 };
struct ZFill {
  uint32_t off16;    // destination offset in the arena, 16-byte units
  uint32_t n16;      // length in 16-byte units
  uint32_t kind;
  uint32_t value;    // ZF_CONST32: word; ZF_SSE: start (low 10 bits)
};
#define ZPQ_MAX_FILLS 80

// Lane-parallel HCOMP "recipes" (zpq_hcomp.cpp): for programs of the makeConfig
// family every H[i] is a short fold over recent bytes, so lane i evaluates its
// own hash natively instead of the warp interpreting the bytecode.
#define ZPQ_RCP_MAX 16
enum { ZR_END = 0, ZR_FOLD, ZR_SHL8ADD, ZR_SHR, ZR_SHL, ZR_ROLL, ZR_WORD, ZR_WORD2 };
enum { ZS_NONE = 0, ZS_BYTE, ZS_POSAND, ZS_POSMOD, ZS_DIST };
enum { ZB_ZERO = 0, ZB_OWN, ZB_LANE };
// op word: op | src<<4 | arg<<8 | off<<16
#define ZR_WORD_OF(op, src, arg, off) ((uint32_t)(op) | (uint32_t)(src) << 4 | (uint32_t)(arg) << 8 | (uint32_t)(off) << 16)
struct ZRecipes {
  uint32_t ok;             // 1: recipes reproduce the HCOMP program exactly (host-verified)
  uint32_t nrounds;
  uint32_t needs_lastpos;  // some recipe reads the H[255+k] "last position of byte k" slots
  uint32_t has_word;
  uint8_t start[32];       // ZB_*
  uint8_t base[32];        // ZB_LANE: lane whose (new) hash seeds this one
  uint8_t round[32];
  uint8_t len[32];
  uint32_t ops[ZPQ_RCP_MAX][32];
};

// Skewed-pipeline ENCODE plan.  While compressing, the coded bit is input data,
// so no component depends on a later one or on the coder: the COMP chain is a
// feed-forward pipeline.  Lane i runs `depth[i]` BYTES behind the head of the
// stream; values cross lanes only through per-lane delay rings, never through
// a same-bit dependency chain.
#define ZPQ_PIPE_STEPS 8
struct ZPipe {
  uint32_t ok;
  uint32_t dmax;                          // deepest stage, in bytes (<= 15)
  uint32_t nsp;                           // special steps: MIX / MIX2 / SSE / AVG in component order
  uint32_t sd[ZPQ_PIPE_STEPS];            // ZSTEP_PACK words (same layout as ZFast.sd)
  uint32_t delay[ZPQ_MAX_LANES];          // 4 bits per special step: depth[owner] - depth[lane] for feeding lanes
  uint8_t depth[ZPQ_MAX_LANES];
  uint8_t dist_byte[ZPQ_MAX_LANES];       // byte value watched by the lane's ZS_DIST op (if has_dist)
  uint8_t has_dist[ZPQ_MAX_LANES];
};

// Wavefront encode plan (zpq_kernels5.cu): one THREAD per (block, component); a component at depth d
// codes chunk w-d in wave w.  Context recipes are flattened so that every thread is self-contained:
// a hash seeded by another component carries a private copy of that component's state.
#define ZPQ_FLAT_MAX 40
#define ZPQ_FLAT_SEGS 4
#define ZR_SEG 8                 // op word: ZR_SEG | start<<4 ; begins the next segment of a flattened recipe
struct ZWave {
  uint32_t ok;
  uint32_t dcoder;                         // depth of the coder stage = depth[n-1] + 1
  uint32_t flat_len[ZPQ_MAX_LANES];
  uint8_t fold_n[ZPQ_MAX_LANES];           // n+1 when the recipe is exactly "hash of the last n bytes" (n = 0..8: zero start, then
                                           // (acc + byte[k] + 512) * 773 for k = 0..n-1): evaluated from the history register
                                           // without interpreting the op list; 0 = general recipe
  alignas(16) uint32_t flat[ZPQ_MAX_LANES][ZPQ_FLAT_MAX];   // rows of 160 bytes, fetched 16 bytes at a time
};

// Slot-scheduled context hashes (two-warp decoder, zpq_kernels6.cu).  In the decoder a lane is a component, so the
// flattened recipes of a warp's lanes are DIFFERENT programs and interpreting them lane by lane diverges on every op.
// The host merges the op lists of all lanes into one sequence of slots whose op kind is the same for every lane
// (a shortest-common-supersequence heuristic, zpq_hcomp.cpp build_hash_plan); a lane that has no op in a slot sits it
// out.  Evaluation is then one uniform loop: slot[s].kind picks the code, w[s][lane] is the lane's operand word,
// slot[s].lanes the participants.  Host-verified against the flat form.
#define ZPQ_HSLOTS 28
// slot kinds: the flat ops, with the two folding ops split by where their operand comes from so that the common case
// (one of the last eight bytes, held in a register) is straight-line code
enum { ZH_SEG = 1, ZH_FOLD_BYTE, ZH_FOLD_OTHER, ZH_SHL8_BYTE, ZH_SHL8_OTHER, ZH_SHR, ZH_SHL, ZH_ROLL, ZH_WORD };
struct ZHashSlot { uint32_t lanes, kind; };   // lanes that have an op in the slot; ZH_*
struct ZHashPlan {
  uint32_t ok, nslots;
  uint8_t last_seg[ZPQ_MAX_LANES];       // index of the lane's last segment (its state receives the final value)
  uint8_t fold_n[ZPQ_MAX_LANES];         // n+1 for lanes whose recipe is the order-n byte hash (ZWave.fold_n): they take no
                                         // slots; all of them are folded together from the history register in one
                                         // predicated 8-step loop
  uint32_t fold_lanes;                   // those lanes
  alignas(8) ZHashSlot slot[ZPQ_HSLOTS];
  uint32_t w[ZPQ_HSLOTS][32];            // operand word per lane (the flat op word; ZH_SEG: segment index in bits 8-9)
  uint32_t w2[ZPQ_HSLOTS][32];           // ZH_WORD: second word (multiplier, first flag)
  // the two-warp decoders split the components over two warps (role 0: ICM ISSE MATCH, role 1: the rest): the
  // slots that touch a lane of each role, so that a warp walks only its own slots (zpq_kernels7.cu)
  uint8_t role_n[2];
  uint8_t role_slots[2][ZPQ_HSLOTS];
  uint32_t role_fold[2];                 // fold lanes of each role
};

// Team plan (round-1 three-warp encoder, removed; the wavefront encoder still reads nmix / mix_comp): the skewed pipeline of ZPipe split over THREE warps per block so that
// the block's instruction stream is issued by three schedulers at once:
//   role 0 = every ICM / ISSE component (one code path, no divergence)
//   role 1 = every MIX component (lane l = input l, cooperative dot product)
//   role 2 = everything else (CM, MATCH, CONST, AVG, MIX2, SSE)
// The arithmetic coder rides with the role that owns the last component.  Predictions cross
// roles (and lanes) through per-component rings of 32 byte-slots x 8 x int16 in the block's arena.
#define ZPQ_ROLES 3
struct ZTeam {
  uint32_t ok;
  uint32_t coder_role;
  uint32_t pmask[ZPQ_ROLES];             // roles this role reads predictions from
  uint32_t cmask[ZPQ_ROLES];             // roles that read this role's predictions
  uint32_t nmix;
  uint32_t mix_comp[ZPQ_FAST_MIX];
  uint32_t nsse;
  uint32_t sse_comp[ZPQ_FAST_SSE];
  uint32_t ring_off16;                   // arena offset of the rings (n x 512 B)
  uint8_t role_of[ZPQ_MAX_LANES];        // component -> role
  uint8_t comp_at[ZPQ_ROLES][32];        // role 0 / 2: lane -> component (0xff: idle lane)
};

// Lock-step decode plan (zpq_kernels8.cu): lane = block, one warp per ICM / ISSE component, one warp for up to four
// CM / MATCH / CONST components, one for the ISSE chains, one for the mixers, one for the scalar steps and the decoders.
#define ZPQ_T_HASH 24
#define ZPQ_T_MISC 4
#define ZPQ_T_SCALAR 4
struct ZTPlan {
  uint32_t ok;
  uint32_t nh;                         // hash warps
  uint32_t nmisc, nisse, nscalar;
  uint8_t hcomp[ZPQ_T_HASH];           // component of hash warp k
  uint8_t mcomp[ZPQ_T_MISC];           // CM / MATCH / CONST components (slots of the misc warp)
  uint8_t icomp[ZPQ_T_HASH];           // ISSE components in component order
  uint8_t ijn[ZPQ_T_HASH];             // their input components
  uint8_t isse_ord[ZPQ_MAX_LANES];     // component -> index in icomp
  uint8_t scomp[ZPQ_T_SCALAR];         // MIX2 / SSE / AVG components in component order
};

struct ZModel {
  uint32_t n, hh, hm;
  uint32_t prog_off;      // offset of the guarded program image in the program pool
  uint32_t prog_len;      // HCOMP bytes (incl. END) without guards
  uint32_t nsteps, nfills;
  uint32_t m_off16, h_off16, r_off16;
  uint32_t m_mask, h_mask;
  uint64_t arena_bytes;   // multiple of 256
  // models with more than 32 components (up to the format's 255): coded by the wide kernel, one thread per block; their
  // component and fill lists do not fit the fixed arrays below and follow the program image in the program pool
  uint32_t wide;          // 1: wide model
  uint32_t wide_comp_off; // byte offset from prog_off of ZComp[n]
  uint32_t wide_fill_off; // byte offset from prog_off of ZFill[nfills]
  uint32_t wst_off16;     // arena: per-component dynamic state, 8 words each (zero filled)
  ZComp comp[ZPQ_MAX_LANES];
  ZStep steps[ZPQ_MAX_STEPS];
  ZFill fills[ZPQ_MAX_FILLS];
  // ---- fast kernel plan (zpq_kernels2.cu) ----
  ZFast fast;
  // ---- pipeline depths of the components (wavefront encoder; dist_byte is read by the decoders too) ----
  ZPipe pipe;
  // ---- role split / mixer list (wavefront encoder's MIX warps) ----
  ZTeam team;
  // ---- wavefront encode plan (zpq_kernels5.cu) ----
  ZWave wave;
  ZRecipes rcp;
  // ---- slot-scheduled hashes for the two-warp decoder (zpq_kernels6.cu) ----
  ZHashPlan hplan;
  // ---- lock-step decoder (zpq_kernels8.cu) ----
  ZTPlan tplan;
};

struct ZBlock {
  uint32_t model;
  uint32_t seg_first, nseg;
  uint32_t pad;
  uint64_t arena_off;
  uint64_t in_off;        // into the batch input buffer
  uint64_t out_off, out_cap;
};

// Model-independent tables (device copies live in the context; the coding
// kernel stages them in shared memory).
struct ZTables {
  int16_t stretch_hi[16384];   // stretcht[16384..32767]; lower half by antisymmetry
  uint16_t squash[4096];
  int32_t dt[1024];
  uint32_t ns[256];            // per state: next0 | next1<<8 | n0<<16 | n1<<24
  uint16_t dt2k[256];
  uint32_t icm_init[256];      // st.cminit(j)
  uint32_t isse_init[512];     // {1<<15, clamp512k(stretch(cminit(j)>>8)*1024)}
};

struct ZKernelArgs {
  const ZModel* models;
  const ZBlock* blocks;
  const uint8_t* progs;
  const uint64_t* seg_len;
  uint64_t* seg_out;
  const uint8_t* in;
  uint8_t* out;
  uint8_t* arena;
  int32_t* status;
  uint64_t* out_len;
  unsigned int* counter;
  const ZTables* tables;
  int nblk;
};


// Arguments of the wavefront encode kernel (zpq_kernels5.cu), one launch per wave.
struct ZWaveArgs {
  ZKernelArgs A;                 // blocks of this launch class, grouped: 32 consecutive blocks of one model = a group
  uint4* pbuf;                   // prediction streams [group][component slot][16 chunk slots][chunk byte][lane < gsize]
  uint32_t* cstate;              // [block][component slot][16 words] state carried between waves (zeroed per batch)
  const uint32_t* group_first;   // first block index of each group
  const uint32_t* group_count;   // blocks in the group (<= 32)
  const uint32_t* blk_group;     // block -> group
  const uint32_t* blk_lane;      // block -> lane inside its group
  int wave, chunk_bytes, ncs, ngroups;   // ncs = component slots per block = max components + 1 (coder)
  int gsize;                             // blocks per group (= active lanes of a component warp), <= 32
  unsigned long long* prof;              // optional [32][2]: cycles and warp count per role (ZPQ_WAVE_PROF=1), else null
  int hash_loop;                         // 1: ICM / ISSE component threads run the nibble-unrolled loop (zpq_kernels5.cu)
  int mix_batch;                         // 1: MIX warps mix the eight bits of a byte together (full-mask mixers; zpq_kernels5.cu)
  int role_major;                        // 1: warp index = component * ngroups + group (role-homogeneous CTAs), 0: group * ncs + component
  int mix_threads;                       // 1: every MIX of the launch has a full c8 mask and >= 256 weight sets and is coded
                                         // by a component thread (eight independent rows per byte); 0: one warp per (block, MIX)
};

#endif
