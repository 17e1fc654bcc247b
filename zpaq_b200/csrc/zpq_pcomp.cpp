// zpq_pcomp.cpp -- SURVEY.md 8(f)-4: the post-processor programs that go with the forward transforms of zpq_bwt.cu.
//
// A block whose payload was transformed before it entered the model (E8E9 address rewriting, BWT) carries the INVERSE
// transform as a ZPAQL program: the header announces the machine (ph, pm), and the program itself is the first thing
// that passes through the model ("1, len_lo, len_hi, bytes..., 0", Compressor::postProcess, libzpaq.cpp:2888-2915).
// For an archive to be byte-identical with libzpaq's, these bytes must equal what makeConfig + Compiler produce for
// the same method (libzpaq.cpp:7135-7296 for BWT, 7296-7323 for E8E9; jump encodings 2575-2662).  They are assembled
// here from opcode numbers by a small structured assembler and pinned by tests against blocks written by the
// unmodified reference (tests/test_prep_transforms.py).
//
// What the programs do (a reader's summary, needed to follow the opcode lists):
//   BWT: bytes 0..255 are appended to M; at end of segment (input 0xffffffff) the last four bytes are the index of the
//   end-of-string slot; bytes are counted into H (complemented indexes), turned into cumulative counts, and a linked
//   list "position -> next position" is threaded through H by walking M in order; the list is followed from the index
//   and every visited byte is written out (or, with E8E9, stored back to M and un-rewritten in a second pass).  Blocks
//   of more than 16 MiB use a traversal that keeps the byte in M instead of packing it into H, and undo E8E9 on the fly
//   through a four-byte window in R4 / R5.
//   E8E9: a five-byte window over the stream; when the oldest byte is E8 / E9 and the newest is 00 / FF, the three
//   bytes in between had the position added by the compressor and get it subtracted again.
#include "zpq_host.h"

namespace zpq {
namespace {

// ZPAQL opcode numbering (opcodelist, libzpaq.cpp:2456-2491): three regular groups
enum Reg { rA = 0, rB = 1, rC = 2, rD = 3, pB = 4, pC = 5, pD = 6, IMM = 7 };
enum Alu { ADD = 0, SUB = 1, MUL = 2, DIV = 3, MOD = 4, AND = 5, ANDN = 6, OR = 7, XOR = 8, SHL = 9, SHR = 10, EQ = 11, LT = 12, GT = 13 };
constexpr int swap_a(Reg r) { return 8 * r; }           // r<>a
constexpr int inc(Reg r) { return 8 * r + 1; }
constexpr int dec(Reg r) { return 8 * r + 2; }
constexpr int inv(Reg r) { return 8 * r + 3; }           // r!
constexpr int zero(Reg r) { return 8 * r + 4; }
constexpr int load_r(Reg r) { return 8 * r + 7; }        // a/b/c/d = r N   (r in rA..rD)
constexpr int set(Reg dst, Reg src) { return 64 + 8 * dst + src; }
constexpr int alu(Alu k, Reg src) { return 128 + 8 * k + src; }
enum { JT = 39, JF = 47, STORE_R = 55, HALT = 56, OUT = 57, JMP = 63, LJ = 255 };

// Structured assembler: the control words of the Compiler, with its encodings.  Short conditional / unconditional jumps
// are two bytes with a signed displacement from the byte after the operand; the "long" forms are a two-byte skip over a
// three-byte absolute jump.
struct PAsm {
  std::vector<uint8_t> code;
  std::vector<size_t> ifs, dos;                 // operand position of the open forward jump; start of the open loop
  bool ok = true;

  PAsm& o(int op) { code.push_back((uint8_t)op); return *this; }
  PAsm& o(int op, int n) { code.push_back((uint8_t)op); code.push_back((uint8_t)n); return *this; }
  PAsm& seq(std::initializer_list<int> ops) { for (int v : ops) code.push_back((uint8_t)v); return *this; }   // opcodes without operands
  bool long_at(size_t a) const { return code[a - 1] == LJ; }

  PAsm& If() { o(JF, 0); ifs.push_back(code.size() - 1); return *this; }
  PAsm& IfNot() { o(JT, 0); ifs.push_back(code.size() - 1); return *this; }
  PAsm& IfL() { o(JT, 3); o(LJ, 0); code.push_back(0); ifs.push_back(code.size() - 2); return *this; }
  PAsm& IfNotL() { o(JF, 3); o(LJ, 0); code.push_back(0); ifs.push_back(code.size() - 2); return *this; }
  PAsm& else_common(bool far) {
    if (ifs.empty()) { ok = false; return *this; }
    const size_t a = ifs.back(); ifs.pop_back();
    const size_t after = code.size() + (far ? 3 : 2);                // first byte of the else branch
    if (long_at(a)) { code[a] = (uint8_t)(after & 255); code[a + 1] = (uint8_t)(after >> 8); }
    else { const size_t j = after - (a + 1); if (j > 127) ok = false; code[a] = (uint8_t)j; }
    if (far) { o(LJ, 0); code.push_back(0); ifs.push_back(code.size() - 2); }
    else { o(JMP, 0); ifs.push_back(code.size() - 1); }
    return *this;
  }
  PAsm& Else() { return else_common(false); }
  PAsm& ElseL() { return else_common(true); }
  PAsm& EndIf() {
    if (ifs.empty()) { ok = false; return *this; }
    const size_t a = ifs.back(); ifs.pop_back();
    if (long_at(a)) { code[a] = (uint8_t)(code.size() & 255); code[a + 1] = (uint8_t)(code.size() >> 8); }
    else { const size_t j = code.size() - a - 1; if (j > 127) ok = false; code[a] = (uint8_t)j; }
    return *this;
  }
  PAsm& Do() { dos.push_back(code.size()); return *this; }
  PAsm& loop_end(int short_op, int skip_op) {      // skip_op: the inverse condition that hops over a long jump (0: none)
    if (dos.empty()) { ok = false; return *this; }
    const size_t a = dos.back(); dos.pop_back();
    const long j = (long)a - (long)code.size() - 2;
    if (j >= -127) return o(short_op, (int)(j & 255));
    if (skip_op) o(skip_op, 3);
    o(LJ, (int)(a & 255)); code.push_back((uint8_t)(a >> 8));
    return *this;
  }
  PAsm& Forever() { return loop_end(JMP, 0); }
  PAsm& While() { return loop_end(JT, JF); }
  PAsm& Until() { return loop_end(JF, JT); }
};

// a = little-endian 24 bits at b-1, b-2, b-3 after "b-- a=*b": the shared "pull one more byte into a" step
void pull_byte(PAsm& p) { p.o(dec(rB)).o(alu(SHL, IMM), 8).o(alu(ADD, pB)); }

// un-rewrite E8 / E9 operands in M[0..d) and write M out (second pass of the BWT + E8E9 blocks up to 16 MiB, and of LZ77 + E8E9)
void e8e9_pass_over_m(PAsm& p, bool end_in_r4 = false) {
  if (end_in_r4) p.o(zero(rB)).o(load_r(rD), 4);
  else p.o(set(rD, rB)).o(zero(rB));
  p.Do();
  p.o(set(rA, rB)).o(alu(EQ, rD)).IfNot();
  p.o(alu(ADD, IMM), 4).o(alu(LT, rD)).If();
  p.o(set(rA, pB)).o(alu(AND, IMM), 254).o(alu(EQ, IMM), 232).If();
  p.o(set(rC, rB)).seq({inc(rB), inc(rB), inc(rB), inc(rB)}).o(set(rA, pB)).o(inc(rA)).o(alu(AND, IMM), 254).o(alu(EQ, IMM), 0).If();
  p.o(dec(rB)).o(set(rA, pB));
  pull_byte(p); pull_byte(p);
  p.o(alu(SUB, rB)).o(inc(rA));
  p.o(set(pB, rA)).o(alu(SHR, IMM), 8).o(inc(rB));
  p.o(set(pB, rA)).o(alu(SHR, IMM), 8).o(inc(rB));
  p.o(set(pB, rA)).o(inc(rB));
  p.EndIf();
  p.o(set(rB, rC));
  p.EndIf();
  p.EndIf();
  p.o(set(rA, pB)).o(OUT).o(inc(rB));
  p.Forever();
  p.EndIf();
}

// thread H[0..c) (loop variable b, bound in c) into the list: for each position, the slot its byte's counter points at
void link_positions(PAsm& p) {
  p.Do();
  p.o(set(rA, rC)).o(alu(GT, rB)).If();
  p.o(set(rD, pB)).o(inv(rD)).o(inc(pD)).o(set(rD, pD)).o(dec(rD)).o(set(pD, rB));
  p.o(inc(rB)).Forever();
  p.EndIf();
}

std::vector<uint8_t> bwt_program(bool e8, bool upto16m, bool* ok) {
  PAsm p;
  p.o(alu(GT, IMM), 255).IfNot();
  p.o(set(pB, rA)).o(inc(rB));
  p.ElseL();
  // index of the end-of-string slot: the last four bytes, little endian -> c, R1 ; size -> R2
  p.o(dec(rB)).o(set(rA, pB));
  pull_byte(p); pull_byte(p); pull_byte(p);
  p.o(set(rC, rA)).o(STORE_R, 1);
  p.o(set(rA, rB)).o(STORE_R, 2);
  // byte counts in H[~(byte + 1)]
  p.Do();
  p.o(set(rA, rB)).o(alu(GT, IMM), 0).If();
  p.o(dec(rB)).o(set(rA, pB)).o(inc(rA)).o(alu(AND, IMM), 255).o(set(rD, rA)).o(inv(rD)).o(inc(pD));
  p.Forever();
  p.EndIf();
  // cumulative counts
  p.o(zero(rD)).o(inv(rD)).o(set(pD, IMM), 1).o(zero(rA));
  p.Do();
  p.o(alu(ADD, pD)).o(set(pD, rA)).o(dec(rD));
  p.o(swap_a(rD)).o(inv(rA)).o(alu(GT, IMM), 255).o(inv(rA)).o(swap_a(rD)).Until();
  // the list: positions before the index, then the positions after it
  p.o(zero(rB));
  link_positions(p);
  p.o(set(rB, rC)).o(inc(rB)).o(load_r(rC), 2);
  link_positions(p);
  if (upto16m) {
    // byte into the low 8 bits of its list cell
    p.o(zero(rB)).Do();
    p.o(set(rA, rC)).o(alu(GT, rB)).If();
    p.o(set(rD, rB)).o(set(rA, pD)).o(alu(SHL, IMM), 8).o(alu(ADD, pB)).o(set(pD, rA));
    p.o(inc(rB)).Forever();
    p.EndIf();
    // follow the list from the index
    p.o(load_r(rD), 1).o(zero(rB)).Do();
    p.o(set(rA, rD)).o(alu(EQ, IMM), 0).IfNot();
    p.o(set(rA, pD)).o(alu(SHR, IMM), 8).o(set(rD, rA));
    if (e8) p.o(set(pB, pD)).o(inc(rB));
    else p.o(set(rA, pD)).o(OUT);
    p.Forever();
    p.EndIf();
    if (e8) e8e9_pass_over_m(p);
  } else if (e8) {
    p.o(load_r(rA), 2).o(dec(rA)).o(STORE_R, 2);
    p.o(zero(rC)).o(load_r(rD), 1).Do();
    p.o(set(rA, rD)).o(alu(EQ, IMM), 0).IfNot();
    p.o(set(rD, pD));
    // newest byte into the top of R4, the byte that leaves R4 into R5
    p.o(set(rB, rD)).o(set(rA, pB)).o(alu(SHL, IMM), 24).o(set(rB, rA));
    p.o(load_r(rA), 4).o(STORE_R, 5).o(alu(SHR, IMM), 8).o(alu(OR, rB)).o(STORE_R, 4);
    p.o(set(rA, rC)).o(alu(GT, IMM), 3).If();
    p.o(load_r(rA), 5).o(alu(AND, IMM), 254).o(alu(EQ, IMM), 232).If();
    p.o(load_r(rA), 4).o(alu(SHR, IMM), 24).o(set(rB, rA)).o(inc(rA)).o(alu(AND, IMM), 254).o(alu(LT, IMM), 2).If();
    p.o(load_r(rA), 4).o(alu(SUB, rC)).o(alu(ADD, IMM), 4).o(alu(SHL, IMM), 8).o(alu(SHR, IMM), 8);
    p.o(swap_a(rB)).o(alu(SHL, IMM), 24).o(alu(ADD, rB)).o(STORE_R, 4);
    p.EndIf();
    p.EndIf();
    p.EndIf();
    p.o(set(rA, rC)).o(alu(GT, IMM), 3).If().o(load_r(rA), 5).o(OUT).EndIf().o(inc(rC));
    p.Forever();
    p.EndIf();
    // up to four bytes still in R4
    p.o(load_r(rB), 4);
    for (int left = 3; left >= 0; --left) {
      p.o(set(rA, rC)).o(alu(GT, IMM), left).o(set(rA, rB)).If().o(OUT).EndIf();
      if (left) p.o(alu(SHR, IMM), 8).o(set(rB, rA));
    }
  } else {
    p.o(load_r(rD), 1).Do();
    p.o(set(rA, rD)).o(alu(EQ, IMM), 0).IfNot();
    p.o(set(rD, pD));
    p.o(set(rB, rD)).o(set(rA, pB)).o(OUT);
    p.Forever();
    p.EndIf();
  }
  p.EndIf();
  p.o(HALT);
  *ok = p.ok && p.ifs.empty() && p.dos.empty();
  return p.code;
}

// "a = bits" then shift `k` bits out of the bit buffer c
void drop_bits(PAsm& p, int k) { p.o(set(rA, rC)).o(alu(SHR, IMM), k).o(set(rC, rA)); }

// LZ77 with variable-length codes (LZBuffer level 1).  R1 = state, R2 = length being read, R3 = offset bits to come,
// R4 = write pointer into M, R5 = low offset bits (blocks above 16 MiB: rb of them), c = bit buffer, d = bits in it.
// States: 0 new code, 1 match length, 5 low offset bits, 2 offset, 3 literal length, 4 literals.
std::vector<uint8_t> lz77_bits_program(bool e8, int rb, bool* ok) {
  PAsm p;
  p.o(alu(GT, IMM), 255).If();
  if (e8) e8e9_pass_over_m(p, true);
  p.o(zero(rA)).o(zero(rB)).o(zero(rC)).o(zero(rD)).o(STORE_R, 1).o(STORE_R, 2).o(STORE_R, 3).o(STORE_R, 4);
  p.o(HALT);
  p.EndIf();
  // the new byte on top of the bit buffer
  p.o(alu(SHL, rD)).o(alu(ADD, rC)).o(set(rC, rA));
  p.o(set(rA, IMM), 8).o(alu(ADD, rD)).o(set(rD, rA));
  // state 0: two bits say literal run (00) or match (offset width class), three more the width inside the class
  p.o(load_r(rA), 1).o(alu(EQ, IMM), 0).If();
  p.o(set(rA, IMM), 1).o(STORE_R, 2);
  p.o(set(rA, rC)).o(alu(AND, IMM), 3).o(alu(GT, IMM), 0).If();
  p.o(dec(rA)).o(alu(SHL, IMM), 3).o(STORE_R, 3);
  drop_bits(p, 2);
  p.o(load_r(rB), 3).o(alu(AND, IMM), 7).o(alu(ADD, rB)).o(STORE_R, 3);
  drop_bits(p, 3);
  p.o(set(rA, rD)).o(alu(SUB, IMM), 5).o(set(rD, rA));
  p.o(set(rA, IMM), 1).o(STORE_R, 1);
  p.Else();
  drop_bits(p, 2);
  p.o(dec(rD)).o(dec(rD));
  p.o(set(rA, IMM), 3).o(STORE_R, 1);
  p.EndIf();
  p.EndIf();
  // state 1: interleaved Elias-gamma length, then its two low bits
  p.Do().o(load_r(rA), 1).o(alu(EQ, IMM), 1).If().o(set(rA, rD)).o(alu(GT, IMM), 2).If();
  p.o(set(rA, rC)).o(alu(AND, IMM), 1).o(alu(EQ, IMM), 1).If();
  drop_bits(p, 1);
  p.o(load_r(rB), 2).o(set(rA, rC)).o(alu(AND, IMM), 1).o(alu(ADD, rB)).o(alu(ADD, rB)).o(STORE_R, 2);
  drop_bits(p, 1);
  p.o(dec(rD)).o(dec(rD));
  p.Else();
  drop_bits(p, 1);
  p.o(load_r(rA), 2).o(alu(SHL, IMM), 2).o(set(rB, rA));
  p.o(set(rA, rC)).o(alu(AND, IMM), 3).o(alu(ADD, rB)).o(STORE_R, 2);
  drop_bits(p, 2);
  p.o(dec(rD)).o(dec(rD)).o(dec(rD));
  p.o(set(rA, IMM), rb ? 5 : 2).o(STORE_R, 1);
  p.EndIf();
  p.Forever().EndIf().EndIf();
  if (rb) {
    p.o(load_r(rA), 1).o(alu(EQ, IMM), 5).If().o(set(rA, rD)).o(alu(GT, IMM), rb - 1).If();
    p.o(set(rA, rC)).o(alu(AND, IMM), (1 << rb) - 1).o(STORE_R, 5);
    drop_bits(p, rb);
    p.o(set(rA, rD)).o(alu(SUB, IMM), rb).o(set(rD, rA));
    p.o(set(rA, IMM), 2).o(STORE_R, 1);
    p.EndIf().EndIf();
  }
  // state 2: the offset, then the copy
  p.o(load_r(rA), 1).o(alu(EQ, IMM), 2).If().o(load_r(rA), 3).o(alu(GT, rD)).IfNot();
  p.o(set(rA, rC)).o(STORE_R, 6).o(set(rA, rD)).o(STORE_R, 7);
  p.o(load_r(rB), 3).o(set(rA, IMM), 1).o(alu(SHL, rB)).o(set(rD, rA));
  p.o(dec(rA)).o(alu(AND, rC)).o(alu(ADD, rD));
  if (rb) p.o(alu(SHL, IMM), rb).o(load_r(rD), 5).o(alu(ADD, rD)).o(alu(SUB, IMM), (1 << rb) - 1);
  p.o(set(rD, rA)).o(load_r(rB), 4).o(set(rA, rB)).o(alu(SUB, rD)).o(set(rC, rA));
  p.o(load_r(rD), 2).Do().o(set(rA, rD)).o(alu(GT, IMM), 0).If().o(dec(rD));
  p.o(set(rA, pC)).o(set(pB, rA)).o(inc(rC)).o(inc(rB));
  if (!e8) p.o(OUT);
  p.Forever().EndIf();
  p.o(set(rA, rB)).o(STORE_R, 4);
  p.o(load_r(rA), 6).o(load_r(rB), 3).o(alu(SHR, rB)).o(set(rC, rA));
  p.o(load_r(rA), 7).o(alu(SUB, rB)).o(set(rD, rA));
  p.o(zero(rA)).o(STORE_R, 1);
  p.EndIf().EndIf();
  // state 3: literal run length
  p.Do().o(load_r(rA), 1).o(alu(EQ, IMM), 3).If().o(set(rA, rD)).o(alu(GT, IMM), 1).If();
  p.o(set(rA, rC)).o(alu(AND, IMM), 1).o(alu(EQ, IMM), 1).If();
  drop_bits(p, 1);
  p.o(load_r(rB), 2).o(alu(AND, IMM), 1).o(alu(ADD, rB)).o(alu(ADD, rB)).o(STORE_R, 2);
  drop_bits(p, 1);
  p.o(dec(rD)).o(dec(rD));
  p.Else();
  drop_bits(p, 1);
  p.o(dec(rD));
  p.o(set(rA, IMM), 4).o(STORE_R, 1);
  p.EndIf();
  p.Forever().EndIf().EndIf();
  // state 4: one literal
  p.o(load_r(rA), 1).o(alu(EQ, IMM), 4).If().o(set(rA, rD)).o(alu(GT, IMM), 7).If();
  p.o(load_r(rB), 4).o(set(rA, rC)).o(set(pB, rA));
  if (!e8) p.o(OUT);
  p.o(inc(rB)).o(set(rA, rB)).o(STORE_R, 4);
  drop_bits(p, 8);
  p.o(set(rA, rD)).o(alu(SUB, IMM), 8).o(set(rD, rA));
  p.o(load_r(rA), 2).o(dec(rA)).o(STORE_R, 2).o(alu(EQ, IMM), 0).If();
  p.o(zero(rA)).o(STORE_R, 1);
  p.EndIf();
  p.EndIf().EndIf();
  p.o(HALT);
  *ok = p.ok && p.ifs.empty() && p.dos.empty();
  return p.code;
}

// Byte-aligned LZ77 (LZBuffer level 2).  d = state (0 new code, 1 literals, 2 last offset byte, 3.. offset bytes to come),
// R1 = length, R2 = offset so far, M = output, b = write pointer.
std::vector<uint8_t> lz77_bytes_program(bool e8, int min_match, bool* ok) {
  PAsm p;
  p.o(alu(GT, IMM), 255).If();
  if (e8) e8e9_pass_over_m(p);
  p.o(zero(rB)).o(zero(rC)).o(zero(rD)).o(zero(rA)).o(STORE_R, 1).o(STORE_R, 2);
  p.o(HALT);
  p.EndIf();
  p.o(set(rC, rA)).o(set(rA, rD)).o(alu(EQ, IMM), 0).If();
  p.o(set(rA, rC)).o(alu(SHR, IMM), 6).o(inc(rA)).o(set(rD, rA));
  p.o(alu(EQ, IMM), 1).If();
  p.o(alu(ADD, rC)).o(STORE_R, 1).o(zero(rA)).o(STORE_R, 2);
  p.Else();
  p.o(inc(rD)).o(set(rA, rC)).o(alu(AND, IMM), 63).o(alu(ADD, IMM), min_match).o(STORE_R, 1).o(zero(rA)).o(STORE_R, 2);
  p.EndIf();
  p.Else();
  p.o(alu(EQ, IMM), 1).If();
  p.o(set(rA, rC)).o(set(pB, rA)).o(inc(rB));
  if (!e8) p.o(OUT);
  p.o(load_r(rA), 1).o(dec(rA)).o(alu(EQ, IMM), 0).If().o(zero(rD)).EndIf().o(STORE_R, 1);
  p.Else();
  p.o(alu(GT, IMM), 2).If();
  p.o(load_r(rA), 2).o(alu(SHL, IMM), 8).o(alu(OR, rC)).o(STORE_R, 2).o(dec(rD));
  p.Else();
  p.o(load_r(rA), 2).o(alu(SHL, IMM), 8).o(alu(OR, rC)).o(set(rC, rA)).o(set(rA, rB)).o(alu(SUB, rC)).o(dec(rA)).o(set(rC, rA));
  p.o(load_r(rD), 1);
  p.Do();
  p.o(set(rA, pC)).o(set(pB, rA)).o(inc(rC)).o(inc(rB));
  if (!e8) p.o(OUT);
  p.o(dec(rD)).o(set(rA, rD)).o(alu(GT, IMM), 0).While();
  p.EndIf();
  p.EndIf();
  p.EndIf();
  p.o(HALT);
  *ok = p.ok && p.ifs.empty() && p.dos.empty();
  return p.code;
}

std::vector<uint8_t> e8e9_program(bool* ok) {
  PAsm p;
  p.o(alu(GT, IMM), 255).If();
  // end of segment: flush the window (c bytes of it, at most four)
  p.o(set(rA, rC)).o(alu(GT, IMM), 4).If();
  p.o(set(rC, IMM), 4);
  p.Else();
  p.o(inv(rA)).o(alu(ADD, IMM), 5).o(alu(SHL, IMM), 3).o(set(rD, rA)).o(set(rA, rB)).o(alu(SHR, rD)).o(set(rB, rA));
  p.EndIf();
  p.Do().o(set(rA, rC)).o(alu(GT, IMM), 0).If();
  p.o(set(rA, rB)).o(OUT).o(alu(SHR, IMM), 8).o(set(rB, rA)).o(dec(rC));
  p.Forever().EndIf();
  p.Else();
  // shift the byte into the window b (newest on top); M[b] remembers the byte that leaves
  p.o(set(pB, rB)).o(alu(SHL, IMM), 24).o(set(rD, rA)).o(set(rA, rB)).o(alu(SHR, IMM), 8).o(alu(ADD, rD)).o(set(rB, rA)).o(inc(rC));
  p.o(set(rA, rC)).o(alu(GT, IMM), 4).If();
  p.o(set(rA, pB)).o(OUT);
  p.o(alu(AND, IMM), 254).o(alu(EQ, IMM), 232).If();
  p.o(set(rA, rB)).o(alu(SHR, IMM), 24).o(inc(rA)).o(alu(AND, IMM), 254).o(alu(EQ, IMM), 0).If();
  p.o(set(rA, rB)).o(alu(SHR, IMM), 24).o(alu(SHL, IMM), 24).o(set(rD, rA));
  p.o(set(rA, rB)).o(alu(SUB, rC)).o(alu(ADD, IMM), 5);
  p.o(alu(SHL, IMM), 8).o(alu(SHR, IMM), 8).o(alu(OR, rD)).o(set(rB, rA));
  p.EndIf();
  p.EndIf();
  p.EndIf();
  p.EndIf();
  p.o(HALT);
  *ok = p.ok && p.ifs.empty() && p.dos.empty();
  return p.code;
}

}  // namespace

// kind = makeConfig's second number: 1 LZ77 with variable-length codes, 2 byte-aligned LZ77 (min_match = its third number),
// 3 BWT, 4 E8E9 only, 5 / 6 / 7 = E8E9 + 1 / 2 / 3.  Returns the program bytes WITHOUT the closing END byte (the prefix adds it).
int pcomp_program(int kind, int size_class, int min_match, std::vector<uint8_t>* code) {
  bool ok = false;
  if (kind == 4) *code = e8e9_program(&ok);
  else if (kind == 3 || kind == 7) *code = bwt_program(kind == 7, size_class <= 4, &ok);
  else if (kind == 1 || kind == 5) *code = lz77_bits_program(kind == 5, size_class > 4 ? size_class - 4 : 0, &ok);
  else if (kind == 2 || kind == 6) *code = lz77_bytes_program(kind == 6, min_match, &ok);
  else return ZPQ_E_METHOD;
  return ok ? ZPQ_OK : ZPQ_E_METHOD;
}

// "1, len lo, len hi, program, END" -- what goes through the model in front of the transformed payload
void pcomp_prefix(const std::vector<uint8_t>& code, std::vector<uint8_t>* prefix) {
  const size_t len = code.size() + 1;
  prefix->clear();
  prefix->push_back(1);
  prefix->push_back((uint8_t)(len & 255));
  prefix->push_back((uint8_t)(len >> 8));
  prefix->insert(prefix->end(), code.begin(), code.end());
  prefix->push_back(0);
}

}  // namespace zpq
