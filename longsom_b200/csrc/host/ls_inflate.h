// Raw DEFLATE (RFC 1951) decoder for BGZF members -- host side of the streaming BAM readers.
//
// A BGZF member is an independent deflate stream of at most 64 KiB whose compressed and inflated sizes are both known
// before decoding starts (BSIZE in the gzip extra field, ISIZE in the trailer), so the decoder works on two whole
// buffers: no streaming state, no window copies, a 64-bit bit buffer refilled eight bytes at a time, two-level
// canonical-Huffman tables (11 bits for literals / lengths, 8 bits for distances) and word-wide match copies.  It is
// what keeps the 16 host cores of a GPU box ahead of the device in the drop-in CLIs (the decode stage of
// BaseCellCounter is inflate-bound: base qualities are close to incompressible, i.e. literal-heavy streams).
//
// Any stream this decoder does not accept (it is strict: exact output size, complete codes) is handed to zlib by the
// callers, so a malformed member still ends in zlib's verdict.
#pragma once
#include <cstdint>
#include <cstring>

namespace lsinf {

constexpr int LBITS = 11, DBITS = 8;
constexpr uint32_t K_LIT = 0u << 13, K_BASE = 1u << 13, K_EOB = 2u << 13, K_LINK = 3u << 13, K_LIT2 = 4u << 13, K_BAD = 7u << 13,
                   K_MASK = 7u << 13;

// entry: bits 0-7 codeword bits to consume (link: size of the subtable in bits), 8-12 extra bits, 13-15 kind, 16-31 value
// (K_LIT2: two literals whose codes together fit the first-level index: first byte in bits 16-23, second in 24-31)
struct Tables {
  uint32_t lit[(1 << LBITS) + 288 * 16];
  uint32_t dist[(1 << DBITS) + 32 * 128];
  uint8_t sub_bits[1 << LBITS];
};

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline uint32_t symbol_entry(bool is_dist, int sym) {
  if (is_dist) {
    if (sym >= 30) return K_BAD;
    return K_BASE | ((uint32_t)DIST_EXTRA[sym] << 8) | ((uint32_t)DIST_BASE[sym] << 16);
  }
  if (sym < 256) return K_LIT | ((uint32_t)sym << 16);
  if (sym == 256) return K_EOB;
  if (sym >= 286) return K_BAD;
  return K_BASE | ((uint32_t)LEN_EXTRA[sym - 257] << 8) | ((uint32_t)LEN_BASE[sym - 257] << 16);
}

// Canonical code of `lens[0..n)` (0 = unused symbol) -> two-level table.  False for an over-subscribed or incomplete
// code (one exception, as in zlib: a distance code with a single symbol of length 1, or no symbol at all).
inline bool build_table(const uint8_t *lens, int n, uint32_t *tab, int tbits, int cap, uint8_t *sub_bits, bool is_dist) {
  int count[16] = {0};
  for (int i = 0; i < n; ++i) ++count[lens[i]];
  count[0] = 0;
  int left = 1, used = 0;
  for (int l = 1; l <= 15; ++l) {
    left = (left << 1) - count[l];
    if (left < 0) return false;
    used += count[l];
  }
  if (left > 0 && !(is_dist && used <= 1)) return false;
  uint32_t next[16];
  {
    uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) {
      code = (code + (uint32_t)count[l - 1]) << 1;
      next[l] = code;
    }
  }
  const int tsize = 1 << tbits;
  for (int i = 0; i < tsize; ++i) tab[i] = K_BAD;
  for (int i = 0; i < tsize; ++i) sub_bits[i] = 0;
  uint16_t rcode[288];
  for (int s = 0; s < n; ++s) {
    const int l = lens[s];
    if (!l) continue;
    uint32_t c = next[l]++, r = 0;
    for (int b = 0; b < l; ++b) r |= ((c >> b) & 1u) << (l - 1 - b);  // deflate packs Huffman codes MSB first
    rcode[s] = (uint16_t)r;
    if (l > tbits) {
      const uint32_t p = r & (uint32_t)(tsize - 1);
      if (l - tbits > sub_bits[p]) sub_bits[p] = (uint8_t)(l - tbits);
    }
  }
  int top = tsize;
  for (int p = 0; p < tsize; ++p) {
    if (!sub_bits[p]) continue;
    const int sz = 1 << sub_bits[p];
    if (top + sz > cap) return false;
    tab[p] = K_LINK | (uint32_t)sub_bits[p] | ((uint32_t)top << 16);
    for (int i = 0; i < sz; ++i) tab[top + i] = K_BAD;
    top += sz;
  }
  for (int s = 0; s < n; ++s) {
    const int l = lens[s];
    if (!l) continue;
    const uint32_t e = symbol_entry(is_dist, s);  // K_BAD for the symbols of the fixed code that never occur
    const uint32_t r = rcode[s];
    if (l <= tbits) {
      for (uint32_t i = r; i < (uint32_t)tsize; i += 1u << l) tab[i] = e | (uint32_t)l;
    } else {
      const uint32_t p = r & (uint32_t)(tsize - 1);
      const uint32_t start = tab[p] >> 16, sb = sub_bits[p];
      for (uint32_t i = r >> tbits; i < (1u << sb); i += 1u << (l - tbits)) tab[start + i] = e | (uint32_t)(l - tbits);
    }
  }
  return true;
}

struct Bits {
  const uint8_t *p, *end;
  uint64_t buf = 0;
  int cnt = 0;       // valid bits in buf
  int overrun = 0;   // zero bytes appended past the end of the input
  inline void refill() {
    if (end - p >= 8) {
      uint64_t w;
      memcpy(&w, p, 8);
      buf |= w << cnt;
      p += (63 - cnt) >> 3;
      cnt |= 56;
    } else {
      while (cnt < 56) {  // 56 .. 63 valid bits afterwards, like the word-wide path (a shift by 64 is undefined)
        if (p < end)
          buf |= (uint64_t)*p++ << cnt;
        else
          ++overrun;
        cnt += 8;
      }
    }
  }
  inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1ull)); }
  inline void drop(int n) {
    buf >>= n;
    cnt -= n;
  }
  inline uint32_t take(int n) {
    const uint32_t v = peek(n);
    drop(n);
    return v;
  }
  // every consumed bit was a real one
  inline bool sound() const { return overrun * 8 <= cnt; }
};

inline bool read_dynamic(Bits &b, Tables &t) {
  static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  b.refill();
  const int hlit = (int)b.take(5) + 257, hdist = (int)b.take(5) + 1, hclen = (int)b.take(4) + 4;
  if (hlit > 286 || hdist > 30) return false;
  uint8_t cl[19] = {0};
  for (int i = 0; i < hclen; ++i) {
    if (b.cnt < 3) b.refill();
    cl[ORDER[i]] = (uint8_t)b.take(3);
  }
  // the code-length code is decoded through a small one-level table (codes of at most 7 bits)
  uint32_t pre[128 + 64];
  uint8_t pre_sub[128];
  if (!build_table(cl, 19, pre, 7, 128 + 64, pre_sub, false)) {
    // symbol_entry() treats 0..18 as literals, which is all that is needed here
    return false;
  }
  uint8_t lens[286 + 30 + 138];
  int i = 0;
  const int total = hlit + hdist;
  while (i < total) {
    b.refill();
    const uint32_t e = pre[b.peek(7)];
    if ((e & K_MASK) != K_LIT) return false;
    b.drop((int)(e & 0xffu));
    const int sym = (int)(e >> 16);
    if (sym < 16) {
      lens[i++] = (uint8_t)sym;
    } else {
      int rep;
      uint8_t v = 0;
      if (sym == 16) {
        if (i == 0) return false;
        v = lens[i - 1];
        rep = 3 + (int)b.take(2);
      } else if (sym == 17) {
        rep = 3 + (int)b.take(3);
      } else {
        rep = 11 + (int)b.take(7);
      }
      if (i + rep > total) return false;
      memset(lens + i, v, (size_t)rep);
      i += rep;
    }
  }
  if (lens[256] == 0) return false;  // no end-of-block code
  return build_table(lens, hlit, t.lit, LBITS, (int)(sizeof t.lit / 4), t.sub_bits, false) &&
         build_table(lens + hlit, hdist, t.dist, DBITS, (int)(sizeof t.dist / 4), t.sub_bits, true);
}

inline void build_fixed(Tables &t) {
  uint8_t lens[288 + 32];
  for (int i = 0; i < 144; ++i) lens[i] = 8;
  for (int i = 144; i < 256; ++i) lens[i] = 9;
  for (int i = 256; i < 280; ++i) lens[i] = 7;
  for (int i = 280; i < 288; ++i) lens[i] = 8;
  for (int i = 0; i < 32; ++i) lens[288 + i] = 5;
  // symbols 286, 287 and distance codes 30, 31 take part in the code but never occur: give them "bad" entries
  uint8_t l2[288];
  memcpy(l2, lens, 288);
  build_table(l2, 288, t.lit, LBITS, (int)(sizeof t.lit / 4), t.sub_bits, false);
  build_table(lens + 288, 32, t.dist, DBITS, (int)(sizeof t.dist / 4), t.sub_bits, true);
}

// Base qualities make the streams literal-heavy with 5-8 bit codes: wherever the first-level index holds two whole
// literal codes, one lookup yields both bytes (the serial chain per lookup -- mask, load, shift -- is the decoder's
// critical path).
inline void pair_literals(Tables &t) {
  uint32_t single[1 << LBITS];
  memcpy(single, t.lit, sizeof single);
  for (uint32_t i = 0; i < (1u << LBITS); ++i) {
    const uint32_t e1 = single[i];
    if ((e1 & K_MASK) != K_LIT) continue;
    const uint32_t l1 = e1 & 0xffu;
    const uint32_t e2 = single[i >> l1];  // the bits above the index are unknown: only codes that fit are final
    if ((e2 & K_MASK) != K_LIT || l1 + (e2 & 0xffu) > (uint32_t)LBITS) continue;
    t.lit[i] = K_LIT2 | (l1 + (e2 & 0xffu)) | (e1 & 0x00ff0000u) | ((e2 & 0x00ff0000u) << 8);
  }
}

// The symbol loop of one Huffman block.  FAST: runs while at least 8 input bytes and 10 + 258 + 8 output bytes are
// left, so that refills, literal stores and word-wide match copies need no bounds checks; returns 0 when that margin
// is gone.  !FAST: one checked step (the tail of the member).  1 = end of block, -1 = malformed stream.
template <bool FAST>
inline int decode_symbols(Bits &b, const Tables &t, const uint8_t *out, uint8_t *&o_ref, uint8_t *const oend) {
  uint8_t *o = o_ref;
  for (;;) {
    if (FAST) {
      if (b.end - b.p < 8 || oend - o < 10 + 258 + 8) {
        o_ref = o;
        return 0;
      }
      uint64_t w;
      memcpy(&w, b.p, 8);
      b.buf |= w << b.cnt;
      b.p += (63 - b.cnt) >> 3;
      b.cnt |= 56;
    } else {
      b.refill();
    }
    uint32_t e = t.lit[b.peek(LBITS)];
    if ((e & K_MASK) == K_LIT || (e & K_MASK) == K_LIT2) {
      // up to five first-level lookups per refill (5 x 11 bits <= 56), one or two literals each
      int budget = FAST ? 5 : 1;
      for (;;) {
        b.drop((int)(e & 0xffu));
        if ((e & K_MASK) == K_LIT2) {
          if (!FAST && oend - o < 2) return -1;
          o[0] = (uint8_t)(e >> 16);
          o[1] = (uint8_t)(e >> 24);
          o += 2;
        } else {
          if (!FAST && o >= oend) return -1;
          *o++ = (uint8_t)(e >> 16);
        }
        if (--budget == 0) break;
        e = t.lit[b.peek(LBITS)];
        if ((e & K_MASK) != K_LIT && (e & K_MASK) != K_LIT2) break;
      }
      if (!FAST) {
        o_ref = o;
        return 0;
      }
      continue;
    }
    if ((e & K_MASK) == K_LINK) {
      b.drop(LBITS);
      e = t.lit[(e >> 16) + b.peek((int)(e & 0xffu))];
      if ((e & K_MASK) == K_LIT) {
        b.drop((int)(e & 0xffu));
        if (!FAST && o >= oend) return -1;
        *o++ = (uint8_t)(e >> 16);
        if (!FAST) {
          o_ref = o;
          return 0;
        }
        continue;
      }
    }
    if ((e & K_MASK) == K_EOB) {
      b.drop((int)(e & 0xffu));
      o_ref = o;
      return 1;
    }
    if ((e & K_MASK) != K_BASE) return -1;
    b.drop((int)(e & 0xffu));
    const uint32_t len = (e >> 16) + b.take((int)((e >> 8) & 31u));
    uint32_t d = t.dist[b.peek(DBITS)];
    if ((d & K_MASK) == K_LINK) {
      b.drop(DBITS);
      d = t.dist[(d >> 16) + b.peek((int)(d & 0xffu))];
    }
    if ((d & K_MASK) != K_BASE) return -1;
    b.drop((int)(d & 0xffu));
    const uint32_t dist = (d >> 16) + b.take((int)((d >> 8) & 31u));
    if (dist > (size_t)(o - out) || len > (size_t)(oend - o)) return -1;
    const uint8_t *s = o - dist;
    if (FAST) {
      // the margin covers 258 + 8 bytes: whole words, no tail handling.  Matches are short on average (a dozen bytes
      // in BAM records), so the first sixteen bytes go out without a loop.
      uint8_t *const stop = o + len;
      if (dist >= 8) {
        uint64_t w0, w1;
        memcpy(&w0, s, 8);
        memcpy(o, &w0, 8);
        memcpy(&w1, s + 8, 8);
        memcpy(o + 8, &w1, 8);
        if (len > 16) {
          s += 16;
          o += 16;
          do {
            memcpy(&w0, s, 8);
            memcpy(o, &w0, 8);
            s += 8;
            o += 8;
          } while (o < stop);
        }
      } else {
        // period < 8: the first eight bytes one by one (each may depend on the one before), then words from a
        // multiple of the period that is at least eight bytes back
        for (int i = 0; i < 8; ++i) o[i] = s[i];
        if (len > 8) {
          const uint32_t d2 = dist * ((7u + dist) / dist);
          const uint8_t *s2 = o + 8 - d2;
          o += 8;
          do {
            uint64_t w;
            memcpy(&w, s2, 8);
            memcpy(o, &w, 8);
            s2 += 8;
            o += 8;
          } while (o < stop);
        }
      }
      o = stop;
    } else if (dist >= 8 && (size_t)(oend - o) >= (size_t)len + 8) {
      uint8_t *const stop = o + len;
      do {
        uint64_t w;
        memcpy(&w, s, 8);
        memcpy(o, &w, 8);
        s += 8;
        o += 8;
      } while (o < stop);
      o = stop;
    } else {
      for (uint32_t i = 0; i < len; ++i) o[i] = s[i];
      o += len;
    }
    if (!FAST) {
      o_ref = o;
      return 0;
    }
  }
}

// Inflate exactly `out_len` bytes from the raw deflate stream in[0..in_len).  True iff the stream is well formed, ends
// with its final block, and produces exactly out_len bytes.  `t` is scratch (one per thread).
inline bool inflate_raw(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len, Tables &t) {
  Bits b;
  b.p = in;
  b.end = in + in_len;
  uint8_t *o = out, *const oend = out + out_len;
  bool last = false;
  while (!last) {
    b.refill();
    last = b.take(1) != 0;
    const uint32_t type = b.take(2);
    if (type == 0) {
      b.drop(b.cnt & 7);
      b.refill();
      const uint32_t len = b.take(16), nlen = b.take(16);
      if (len != (~nlen & 0xffffu)) return false;
      const int real = (b.cnt >> 3) - b.overrun;  // whole bytes still buffered go back to the input
      if (real < 0) return false;
      b.p -= real;
      b.buf = 0;
      b.cnt = 0;
      b.overrun = 0;
      if ((size_t)(b.end - b.p) < len || (size_t)(oend - o) < len) return false;
      memcpy(o, b.p, len);
      o += len;
      b.p += len;
      continue;
    }
    if (type == 1) {
      build_fixed(t);
    } else if (type == 2) {
      if (!read_dynamic(b, t)) return false;
    } else {
      return false;
    }
    pair_literals(t);
    for (;;) {
      int r = decode_symbols<true>(b, t, out, o, oend);
      if (r == 0) r = decode_symbols<false>(b, t, out, o, oend);
      if (r < 0) return false;
      if (r == 1) break;
    }
    if (b.cnt < 0 || !b.sound()) return false;
  }
  return o == oend && b.cnt >= 0 && b.sound();
}

}  // namespace lsinf
