// zpq_frame.cpp -- the kept host-side framing around the coded data: SHA-1
// (libzpaq.cpp:101-177 computes the same FIPS 180 digest), block/segment
// prologue (Compressor::writeTag/startBlock/startSegment, libzpaq.cpp:2776-2883)
// and an in-memory block scanner (Decompresser::findBlock/findFilename/
// readComment/readSegmentEnd + Decoder::skip, libzpaq.cpp:2158-2181, 2247-2373).
#include "zpq_host.h"
#include <string.h>

namespace zpq {

// ---------------------------------------------------------------- SHA-1 (FIPS 180-4)
static inline uint32_t rol(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

static void sha1_block(uint32_t h[5], const uint8_t* p) {
  uint32_t w[80];
  for (int i = 0; i < 16; ++i) w[i] = (uint32_t)p[i * 4] << 24 | (uint32_t)p[i * 4 + 1] << 16 | (uint32_t)p[i * 4 + 2] << 8 | p[i * 4 + 3];
  for (int i = 16; i < 80; ++i) w[i] = rol(w[i - 3] ^ w[i - 8] ^ w[i - 14] ^ w[i - 16], 1);
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
  for (int i = 0; i < 80; ++i) {
    uint32_t f, k;
    if (i < 20) { f = (b & c) | (~b & d); k = 0x5A827999u; }
    else if (i < 40) { f = b ^ c ^ d; k = 0x6ED9EBA1u; }
    else if (i < 60) { f = (b & c) | (b & d) | (c & d); k = 0x8F1BBCDCu; }
    else { f = b ^ c ^ d; k = 0xCA62C1D6u; }
    const uint32_t t = rol(a, 5) + f + e + k + w[i];
    e = d; d = c; c = rol(b, 30); b = a; a = t;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e;
}

void Sha1::reset() {
  h[0] = 0x67452301u; h[1] = 0xEFCDAB89u; h[2] = 0x98BADCFEu; h[3] = 0x10325476u; h[4] = 0xC3D2E1F0u;
  len = 0; fill = 0;
}

void Sha1::update(const uint8_t* p, uint64_t n) {
  len += n;
  if (fill) {
    while (n && fill < 64) { buf[fill++] = *p++; --n; }
    if (fill == 64) { sha1_block(h, buf); fill = 0; }
  }
  while (n >= 64) { sha1_block(h, p); p += 64; n -= 64; }
  while (n) { buf[fill++] = *p++; --n; }
}

void Sha1::finish(uint8_t out[20]) {
  const uint64_t bits = len * 8;
  uint8_t pad = 0x80;
  update(&pad, 1);
  pad = 0;
  while (fill != 56) update(&pad, 1);
  uint8_t lb[8];
  for (int i = 0; i < 8; ++i) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
  update(lb, 8);
  for (int i = 0; i < 5; ++i) { out[i * 4] = h[i] >> 24; out[i * 4 + 1] = h[i] >> 16; out[i * 4 + 2] = h[i] >> 8; out[i * 4 + 3] = h[i]; }
  reset();
}

void sha1_of(const uint8_t* p, uint64_t n, uint8_t out[20]) {
  Sha1 s;
  s.update(p, n);
  s.finish(out);
}

// ---------------------------------------------------------------- writing
static const uint8_t kTag[13] = {0x37, 0x6b, 0x53, 0x74, 0xa0, 0x31, 0x83, 0xd3, 0x8c, 0xb2, 0x28, 0xb0, 0xd3};

void write_block_prologue(std::vector<uint8_t>* out, const uint8_t* hdr, int hlen,
                          const char* filename, const char* comment) {
  out->clear();
  out->insert(out->end(), kTag, kTag + 13);             // writeTag, libzpaq.cpp:2776
  out->push_back('z'); out->push_back('P'); out->push_back('Q');
  out->push_back((uint8_t)(1 + (hdr[6] == 0)));         // level 1, or 2 when unmodelled (libzpaq.cpp:2864)
  out->push_back(1);
  out->insert(out->end(), hdr, hdr + hlen);             // ZPAQL::write, libzpaq.cpp:866
  out->push_back(1);                                    // startSegment, libzpaq.cpp:2871
  for (const char* p = filename; p && *p; ++p) out->push_back((uint8_t)*p);
  out->push_back(0);
  for (const char* p = comment; p && *p; ++p) out->push_back((uint8_t)*p);
  out->push_back(0);
  out->push_back(0);                                    // reserved
}

// ---------------------------------------------------------------- scanning
static int bad(std::string* err, const char* msg, int code) { if (err) *err = msg; return code; }

int scan_archive(const uint8_t* a, uint64_t n, std::vector<BlockView>* blocks, std::string* err) {
  blocks->clear();
  uint64_t pos = 0;
  for (;;) {
    // findBlock: locate tag + "zPQ" (the reference uses 4 rolling hashes over the same 16 bytes)
    uint64_t k = pos;
    bool found = false;
    for (; k + 16 <= n; ++k)
      if (a[k] == kTag[0] && memcmp(a + k, kTag, 13) == 0 && a[k + 13] == 'z' && a[k + 14] == 'P' && a[k + 15] == 'Q') { found = true; break; }
    if (!found) break;
    pos = k + 16;
    if (pos + 4 > n) return bad(err, "unexpected end of file", ZPQ_E_EOF);
    BlockView B;
    B.level = a[pos];
    if (B.level != 1 && B.level != 2) return bad(err, "unsupported ZPAQ level", ZPQ_E_HEADER);
    if (a[pos + 1] != 1) return bad(err, "unsupported ZPAQL type", ZPQ_E_HEADER);
    pos += 2;
    const int hsize = a[pos] + 256 * a[pos + 1];
    if (pos + 2 + (uint64_t)hsize > n) return bad(err, "unexpected end of file", ZPQ_E_EOF);
    if (hsize < 6) return bad(err, "invalid block header", ZPQ_E_HEADER);   // hh hm ph pm n ... END: the component count is read below
    B.hdr_off = pos; B.hlen = hsize + 2;
    pos += (uint64_t)hsize + 2;
    for (;;) {
      if (pos >= n) return bad(err, "unexpected end of file", ZPQ_E_EOF);
      if (a[pos] == 255) { ++pos; break; }                       // end of block
      if (a[pos] != 1) return bad(err, "missing segment or end of block", ZPQ_E_HEADER);
      ++pos;
      SegmentView S;
      memset(&S, 0, sizeof(S));
      while (pos < n && a[pos]) ++pos;                           // filename
      if (pos >= n) return bad(err, "unexpected EOF", ZPQ_E_EOF);
      ++pos;
      S.comment_off = pos;
      while (pos < n && a[pos]) ++pos;                           // comment
      if (pos >= n) return bad(err, "unexpected EOF", ZPQ_E_EOF);
      S.comment_len = pos - S.comment_off;
      ++pos;
      if (pos >= n || a[pos] != 0) return bad(err, "missing reserved byte", ZPQ_E_HEADER);
      ++pos;
      // Decoder::skip (modelled): the coded bytes never contain 4 zeros in a row
      // (libzpaq.cpp:2414); the terminator is the LAST 4 zeros before the trailer byte.
      S.coded_off = pos;
      uint64_t e = pos;
      if (a[B.hdr_off + 6] == 0) {
        // unmodelled block: 4-byte big-endian chunk lengths until a zero length (libzpaq.cpp:2168-2180)
        for (;;) {
          if (e + 4 > n) return bad(err, "skipped to EOF", ZPQ_E_EOF);
          const uint64_t len = (uint64_t)a[e] << 24 | (uint64_t)a[e + 1] << 16 | (uint64_t)a[e + 2] << 8 | a[e + 3];
          e += 4;
          if (len == 0) break;
          if (e + len > n) return bad(err, "skipped to EOF", ZPQ_E_EOF);
          e += len;
        }
      } else {
        uint32_t win = 0xffffffffu;
        bool got = false;
        for (; e < n; ++e) { win = win << 8 | a[e]; if (e >= pos + 3 && win == 0) { got = true; break; } }
        if (!got) return bad(err, "unexpected end of file", ZPQ_E_EOF);
        ++e;                                                     // one past the 4th zero
        while (e < n && a[e] == 0) ++e;                          // "might be more than 4" (libzpaq.cpp:2165)
      }
      if (e >= n) return bad(err, "unexpected end of file", ZPQ_E_EOF);
      S.coded_len = e - S.coded_off;
      pos = e;
      if (a[pos] == 253) {
        if (pos + 21 > n) return bad(err, "unexpected end of file", ZPQ_E_EOF);
        S.has_sha1 = true; memcpy(S.sha1, a + pos + 1, 20); pos += 21;
      }
      else if (a[pos] == 254) { S.has_sha1 = false; ++pos; }
      else return bad(err, "missing end of segment marker", ZPQ_E_HEADER);
      B.segs.push_back(S);
    }
    B.end = pos;
    blocks->push_back(B);
  }
  return ZPQ_OK;
}

}  // namespace zpq
