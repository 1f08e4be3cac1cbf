// replay_source.hpp -- the replay front door of the reference: `--load <dir>` makes main.cpp
// read frames from ImageSourceFiles (video.h:24-38: "<dir>/%08d.png" by frame id through cv::imread; main.cpp:446-448,
// :503-519: ids are consecutive and the camera alternates, so frames id and id+2 belong to the same camera), and
// `--save <dir>` writes them in that format (main.cpp:373-398).
//
// This header mirrors that source for the batched replay (sfe_replay_pairs): ImageSourceFiles::GetObservation with
// the reference's signature and failure behaviour (false when the file is missing or undecodable), plus LoadPairs,
// which fills the contiguous host buffers the replay entry takes.  OpenCV is not available to this repository's
// C++, so the PNG files are decoded here: 8-bit non-interlaced PNGs of colour type gray, RGB, palette,
// gray+alpha and RGBA (what cv::imwrite produces for CV_8UC1/3/4), returned as 3-channel BGR like
// cv::imread(IMREAD_COLOR).  Host code only; nothing here touches the GPU.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

namespace sfe {

struct BgrImage {
  int cols = 0, rows = 0;
  std::vector<uint8_t> data;  // rows x cols x 3, BGR, dense
  size_t step() const { return (size_t)cols * 3; }
};

namespace png_detail {

// ---- RFC 1951 inflate (stored, fixed and dynamic Huffman blocks).  The canonical-Huffman decoder (per-length code
// counts plus a symbol table, decoded bit by bit against a running first/index pair) is the scheme of Mark Adler's
// public-domain "puff" reference inflater (zlib contrib/puff), restated for this header.  `limit` bounds the output:
// replay directories are external input, and a small file must not inflate without bound.
struct BitReader {
  const uint8_t* p;
  size_t n, pos = 0;
  uint32_t acc = 0;
  int cnt = 0;
  bool ok = true;
  int bits(int k) {
    while (cnt < k) {
      if (pos >= n) { ok = false; return 0; }
      acc |= (uint32_t)p[pos++] << cnt;
      cnt += 8;
    }
    const int v = (int)(acc & ((1u << k) - 1));
    acc >>= k;
    cnt -= k;
    return v;
  }
};

struct Huffman {
  uint16_t count[16], symbol[288];
  bool build(const uint8_t* len, int n) {
    memset(count, 0, sizeof(count));
    for (int i = 0; i < n; ++i) count[len[i]]++;
    count[0] = 0;
    int left = 1;
    for (int l = 1; l < 16; ++l) {
      left = (left << 1) - count[l];
      if (left < 0) return false;
    }
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + count[l];
    for (int i = 0; i < n; ++i)
      if (len[i]) symbol[offs[len[i]]++] = (uint16_t)i;
    return true;
  }
  int decode(BitReader& br) const {
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; ++l) {
      code |= br.bits(1);
      if (!br.ok) return -1;
      const int c = count[l];
      if (code - c < first) return symbol[index + (code - first)];
      index += c;
      first = (first + c) << 1;
      code <<= 1;
    }
    return -1;
  }
};

inline bool inflate(const uint8_t* src, size_t n, std::vector<uint8_t>& out, size_t limit) {
  static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  static const uint16_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  static const uint16_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  BitReader br{src, n};
  for (;;) {
    const int last = br.bits(1), type = br.bits(2);
    if (!br.ok) return false;
    if (type == 0) {
      br.acc = 0; br.cnt = 0;  // to the byte boundary
      if (br.pos + 4 > n) return false;
      const unsigned len = src[br.pos] | (src[br.pos + 1] << 8), nlen = src[br.pos + 2] | (src[br.pos + 3] << 8);
      br.pos += 4;
      if ((len ^ 0xffffu) != nlen || br.pos + len > n || out.size() + len > limit) return false;
      out.insert(out.end(), src + br.pos, src + br.pos + len);
      br.pos += len;
    } else if (type == 1 || type == 2) {
      Huffman hl, hd;
      uint8_t lens[320];
      if (type == 1) {
        int i = 0;
        for (; i < 144; ++i) lens[i] = 8;
        for (; i < 256; ++i) lens[i] = 9;
        for (; i < 280; ++i) lens[i] = 7;
        for (; i < 288; ++i) lens[i] = 8;
        hl.build(lens, 288);
        for (i = 0; i < 30; ++i) lens[i] = 5;
        hd.build(lens, 30);
      } else {
        const int nlen = br.bits(5) + 257, ndist = br.bits(5) + 1, ncode = br.bits(4) + 4;
        if (!br.ok || nlen > 286 || ndist > 30) return false;
        uint8_t cl[19] = {0};
        for (int i = 0; i < ncode; ++i) cl[order[i]] = (uint8_t)br.bits(3);
        Huffman hc;
        if (!hc.build(cl, 19)) return false;
        int i = 0;
        while (i < nlen + ndist) {
          int sym = hc.decode(br);
          if (sym < 0) return false;
          if (sym < 16) lens[i++] = (uint8_t)sym;
          else {
            int rep, val = 0;
            if (sym == 16) { if (i == 0) return false; val = lens[i - 1]; rep = 3 + br.bits(2); }
            else if (sym == 17) rep = 3 + br.bits(3);
            else rep = 11 + br.bits(7);
            if (!br.ok || i + rep > nlen + ndist) return false;
            while (rep--) lens[i++] = (uint8_t)val;
          }
        }
        if (!hl.build(lens, nlen) || !hd.build(lens + nlen, ndist)) return false;
      }
      for (;;) {
        int sym = hl.decode(br);
        if (sym < 0) return false;
        if (sym < 256) {
          if (out.size() >= limit) return false;
          out.push_back((uint8_t)sym);
        } else if (sym == 256) break;
        else {
          sym -= 257;
          if (sym >= 29) return false;
          const int len = lbase[sym] + br.bits(lext[sym]);
          const int ds = hd.decode(br);
          if (ds < 0 || ds >= 30) return false;
          const size_t dist = dbase[ds] + (size_t)br.bits(dext[ds]);
          if (!br.ok || dist > out.size() || out.size() + (size_t)len > limit) return false;
          size_t from = out.size() - dist;
          for (int k = 0; k < len; ++k) out.push_back(out[from + k]);
        }
      }
    } else {
      return false;
    }
    if (last) return true;
  }
}

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

}  // namespace png_detail

// Decodes an in-memory PNG to BGR (cv::imread(..., IMREAD_COLOR) semantics for 8-bit images).  Returns false on
// anything it does not handle (16-bit, interlaced, corrupt data).
inline bool DecodePng(const uint8_t* file, size_t n, BgrImage* img) {
  using namespace png_detail;
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (n < 8 || memcmp(file, sig, 8) != 0) return false;
  size_t pos = 8;
  uint32_t w = 0, h = 0;
  int depth = 0, ctype = -1, interlace = 0;
  std::vector<uint8_t> idat, plte;
  while (pos + 12 <= n) {
    const uint32_t len = be32(file + pos);
    const uint8_t* type = file + pos + 4;
    const uint8_t* body = file + pos + 8;
    if (pos + 12 + (size_t)len > n) return false;
    if (!memcmp(type, "IHDR", 4) && len >= 13) {
      w = be32(body); h = be32(body + 4);
      depth = body[8]; ctype = body[9]; interlace = body[12];
    } else if (!memcmp(type, "PLTE", 4)) plte.assign(body, body + len);
    else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
    else if (!memcmp(type, "IEND", 4)) break;
    pos += 12 + (size_t)len;
  }
  int ch;
  switch (ctype) { case 0: ch = 1; break; case 2: ch = 3; break; case 3: ch = 1; break; case 4: ch = 2; break; case 6: ch = 4; break; default: return false; }
  if (depth != 8 || interlace != 0 || w == 0 || h == 0 || w > 32768 || h > 32768 || idat.size() < 6) return false;
  std::vector<uint8_t> raw;
  // the decoded size is known from IHDR; deflate expands at most 1032x, so a header that promises more than the
  // IDAT bytes can hold is rejected before anything is allocated, and inflate() stops at the expected size
  const size_t expected = (size_t)h * ((size_t)w * ch + 1);
  if (expected > idat.size() * 1032 + 1024) return false;
  raw.reserve(expected);
  if (!inflate(idat.data() + 2, idat.size() - 2, raw, expected)) return false;  // skip the 2-byte zlib header; Adler-32 not checked
  const size_t stride = (size_t)w * ch;
  if (raw.size() < (size_t)h * (stride + 1)) return false;
  // undo the scanline filters in place (PNG spec 9.2)
  std::vector<uint8_t> zero(stride, 0);
  for (uint32_t y = 0; y < h; ++y) {
    uint8_t* line = raw.data() + (size_t)y * (stride + 1);
    const int f = line[0];
    uint8_t* cur = line + 1;
    const uint8_t* up = y ? raw.data() + (size_t)(y - 1) * (stride + 1) + 1 : zero.data();
    for (size_t i = 0; i < stride; ++i) {
      const int a = i >= (size_t)ch ? cur[i - ch] : 0, b = up[i], c = i >= (size_t)ch ? up[i - ch] : 0;
      int v = cur[i];
      switch (f) { case 0: break; case 1: v += a; break; case 2: v += b; break; case 3: v += (a + b) >> 1; break; case 4: v += paeth(a, b, c); break; default: return false; }
      cur[i] = (uint8_t)v;
    }
  }
  img->cols = (int)w; img->rows = (int)h;
  img->data.resize((size_t)w * h * 3);
  for (uint32_t y = 0; y < h; ++y) {
    const uint8_t* s = raw.data() + (size_t)y * (stride + 1) + 1;
    uint8_t* d = img->data.data() + (size_t)y * w * 3;
    for (uint32_t x = 0; x < w; ++x, s += ch, d += 3) {
      uint8_t r, g, b;
      if (ctype == 0 || ctype == 4) r = g = b = s[0];
      else if (ctype == 3) {
        if ((size_t)s[0] * 3 + 2 >= plte.size()) return false;
        r = plte[s[0] * 3]; g = plte[s[0] * 3 + 1]; b = plte[s[0] * 3 + 2];
      } else { r = s[0]; g = s[1]; b = s[2]; }
      d[0] = b; d[1] = g; d[2] = r;  // cv::imread returns BGR; alpha is dropped
    }
  }
  return true;
}

inline bool ReadFile(const std::string& path, std::vector<uint8_t>* out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  out->resize(sz > 0 ? (size_t)sz : 0);
  const bool ok = sz > 0 && fread(out->data(), 1, (size_t)sz, f) == (size_t)sz;
  fclose(f);
  return ok;
}

// video.h:24-38
class ImageSourceFiles {
 public:
  explicit ImageSourceFiles(const std::string& dir) : dir_(dir) {}
  bool Init() { return true; }
  bool GetObservation(int /*camera*/, int frame_id, BgrImage* img) const {
    char b[100];
    snprintf(b, sizeof(b), "%08d.png", frame_id);
    std::vector<uint8_t> bytes;
    return ReadFile(dir_ + "/" + b, &bytes) && DecodePng(bytes.data(), bytes.size(), img);
  }

  // Consecutive frames first, first+1, ... into ONE contiguous buffer, for sfe_replay_sequence (pair i = (i, i + 2): the camera
  // alternates every frame, main.cpp:503-519).  Stops at the first missing or differently sized frame; returns the number of
  // frames loaded.
  int LoadSequence(int first_id, int max_frames, std::vector<uint8_t>* frames_bgr, int* cols, int* rows) const {
    frames_bgr->clear();
    int n = 0;
    BgrImage a;
    for (; n < max_frames; ++n) {
      if (!GetObservation(0, first_id + n, &a)) break;
      if (n == 0) { *cols = a.cols; *rows = a.rows; }
      if (a.cols != *cols || a.rows != *rows) break;
      frames_bgr->insert(frames_bgr->end(), a.data.begin(), a.data.end());
    }
    return n;
  }

  // Consecutive same-camera pairs (id, id + 2) for ids first, first+1, ... (main.cpp:503-519: the camera alternates
  // every frame) into the contiguous buffers sfe_replay_pairs takes.  Stops at the first missing frame, as the
  // reference's main loop does; returns the number of pairs loaded.
  int LoadPairs(int first_id, int max_pairs, std::vector<uint8_t>* from_bgr, std::vector<uint8_t>* to_bgr, int* cols, int* rows) const {
    from_bgr->clear();
    to_bgr->clear();
    int n = 0;
    BgrImage a, b;
    for (; n < max_pairs; ++n) {
      if (!GetObservation(0, first_id + n, &a) || !GetObservation(0, first_id + n + 2, &b)) break;
      if (n == 0) { *cols = a.cols; *rows = a.rows; }
      if (a.cols != *cols || a.rows != *rows || b.cols != *cols || b.rows != *rows) break;
      from_bgr->insert(from_bgr->end(), a.data.begin(), a.data.end());
      to_bgr->insert(to_bgr->end(), b.data.begin(), b.data.end());
    }
    return n;
  }

 private:
  const std::string dir_;
};

}  // namespace sfe
