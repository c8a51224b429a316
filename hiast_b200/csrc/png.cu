// Pseudo-label PNG writer on the device.
//
// Replaces `cv2.imwrite(path, plbl.astype(np.uint8))` of workflows/pseudo_label_generator.py:43-46: the label maps
// never leave the GPU uncompressed; what crosses PCIe is the finished file (signature, IHDR, IDAT..., IEND), typically
// 30-50x smaller than the 2 MB map.  The stream layout is restated bit for bit by oracle/png.py (see its header for the
// format): Up-filtered rows, 128-byte chunks tokenised independently into literals and distance-1 matches, fixed Huffman
// code, one IDAT chunk per segment of <= 256 chunks closed by a sync flush, stored-block fallback for segments that do not
// compress, Adler-32 of the filtered stream and CRC-32 of every chunk computed on the device.
//
// Three launches per batch of images:
//   k_png_count   one CTA per segment, one thread per chunk, no shared staging: equal-to-left-neighbour masks by SIMD byte
//                 compares, closed-form bit counts, dp4a Adler partial sums
//   k_png_offsets one CTA: segment modes and sizes, chunk offsets, Adler combination, file offsets (packed blob), headers
//   k_png_emit    one CTA per segment: bit offsets by a block scan, tokens OR-ed into a shared-memory bit buffer laid out
//                 with the destination's word alignment, chunk CRC by a shift-and-combine tree, word copy to the blob
#include "common.cuh"

namespace hiast {
namespace {

constexpr int kPngChunk = 128;
constexpr int kPngThreads = 256;  // = chunks per segment at most
constexpr uint32_t kCrcPoly = 0xEDB88320u;
constexpr unsigned long long kAdlerMod = 65521ull;
constexpr uint32_t kFilterUp = 2;
constexpr int kCrcLevels = 8;
constexpr int kCrcMaxPiece = 144;

// 128-bit chunk masks as two words.  (`unsigned __int128` compiles, but nvcc lowers its variable shifts to out-of-line
// library routines: half of the emit kernel's samples had no source line.)
struct u128 {
  unsigned long long lo, hi;
};

struct PngGeom {
  int H, W, cpr, R, S, Lmax;
};

inline bool png_geom(int H, int W, PngGeom* g) {
  if (H < 1 || W < 1) return false;
  const int cpr = (W + kPngChunk - 1) / kPngChunk;
  if (cpr > kPngThreads) return false;
  g->H = H;
  g->W = W;
  g->cpr = cpr;
  g->R = std::min(H, kPngThreads / cpr);
  g->S = (H + g->R - 1) / g->R;
  g->Lmax = g->R * (W + 1);
  return true;
}

struct PngWorkspace {
  uint32_t* seg_bits;
  uint32_t* seg_s1;
  unsigned long long* seg_s2;
  uint32_t* seg_off;
  uint32_t* seg_len;  // bit 31: stored mode
  uint32_t* adler;
  uint16_t* chunk_bits;
};

inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

inline size_t png_carve(void* base, int n_images, const PngGeom& g, PngWorkspace* ws) {
  const size_t ns = static_cast<size_t>(n_images) * g.S;
  size_t off = 0;
  char* p = static_cast<char*>(base);
  auto take = [&](size_t bytes) {
    char* q = p ? p + off : nullptr;
    off += al256(bytes);
    return q;
  };
  char* a = take(ns * 8);
  char* b = take(ns * 4);
  char* c = take(ns * 4);
  char* d = take(ns * 4);
  char* e = take(ns * 4);
  char* f = take(static_cast<size_t>(n_images) * 4);
  char* h = take(ns * kPngThreads * 2);
  if (ws) {
    ws->seg_s2 = reinterpret_cast<unsigned long long*>(a);
    ws->seg_bits = reinterpret_cast<uint32_t*>(b);
    ws->seg_s1 = reinterpret_cast<uint32_t*>(c);
    ws->seg_off = reinterpret_cast<uint32_t*>(d);
    ws->seg_len = reinterpret_cast<uint32_t*>(e);
    ws->adler = reinterpret_cast<uint32_t*>(f);
    ws->chunk_bits = reinterpret_cast<uint16_t*>(h);
  }
  return off;
}

// ---- CRC-32 shift factors: x^(8 * m * 2^level) mod P for piece lengths m < kCrcMaxPiece (reflected representation) ----
constexpr uint32_t multmodp_c(uint32_t a, uint32_t b) {
  uint32_t p = 0;
  for (int i = 0; i < 32; ++i) {
    if (a & (0x80000000u >> i)) p ^= b;
    b = (b & 1u) ? (b >> 1) ^ kCrcPoly : b >> 1;
  }
  return p;
}

struct CrcShiftTable {
  uint32_t v[kCrcLevels][kCrcMaxPiece];
};

constexpr CrcShiftTable make_crc_shift() {
  CrcShiftTable t{};
  t.v[0][0] = 0x80000000u;
  for (int m = 1; m < kCrcMaxPiece; ++m) t.v[0][m] = multmodp_c(t.v[0][m - 1], 0x00800000u);
  for (int l = 1; l < kCrcLevels; ++l)
    for (int m = 0; m < kCrcMaxPiece; ++m) t.v[l][m] = multmodp_c(t.v[l - 1][m], t.v[l - 1][m]);
  return t;
}

__constant__ CrcShiftTable c_crc_shift = make_crc_shift();

__device__ __forceinline__ uint32_t multmodp(uint32_t a, uint32_t b) {
  uint32_t p = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    p ^= (a & (0x80000000u >> i)) ? b : 0u;
    b = (b >> 1) ^ ((b & 1u) ? kCrcPoly : 0u);
  }
  return p;
}

// ---- 128-bit mask helpers ----
__device__ __forceinline__ u128 mk(unsigned long long lo, unsigned long long hi) {
  u128 r;
  r.lo = lo;
  r.hi = hi;
  return r;
}
__device__ __forceinline__ u128 operator&(u128 a, u128 b) { return mk(a.lo & b.lo, a.hi & b.hi); }
__device__ __forceinline__ u128 operator|(u128 a, u128 b) { return mk(a.lo | b.lo, a.hi | b.hi); }
__device__ __forceinline__ u128 operator~(u128 a) { return mk(~a.lo, ~a.hi); }
__device__ __forceinline__ bool is_zero(u128 a) { return (a.lo | a.hi) == 0ull; }
__device__ __forceinline__ u128 shr(u128 a, int s) {  // 0 <= s < 128
  if (s & 64) return mk(a.hi >> (s & 63), 0ull);
  return mk((a.lo >> s) | ((a.hi << 1) << (63 - s)), a.hi >> s);
}
__device__ __forceinline__ u128 shl_small(u128 a, int s) {  // 0 < s < 64
  return mk(a.lo << s, (a.hi << s) | (a.lo >> (64 - s)));
}
__device__ __forceinline__ u128 low_mask(int n) {  // n low bits set, 0 <= n <= 128
  if (n >= 128) return mk(~0ull, ~0ull);
  if (n >= 64) return mk(~0ull, (1ull << (n - 64)) - 1ull);
  return mk((1ull << n) - 1ull, 0ull);
}
__device__ __forceinline__ int ctz128(u128 v) {  // v != 0
  if (v.lo) return __ffsll(static_cast<long long>(v.lo)) - 1;
  return 64 + __ffsll(static_cast<long long>(v.hi)) - 1;
}
__device__ __forceinline__ int popc128(u128 v) { return __popcll(v.lo) + __popcll(v.hi); }
__device__ __forceinline__ u128 make128(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  return mk(static_cast<unsigned long long>(b) << 32 | a, static_cast<unsigned long long>(d) << 32 | c);
}
__device__ __forceinline__ uint32_t byte_mask_to_nibble(uint32_t m) {  // 0xFF/0x00 per byte -> 4 bits
  return (((m & 0x01010101u) * 0x01020408u) >> 24) & 0xFu;
}

// ---- fixed Huffman code (RFC 1951 3.2.6), bit-reversed for the LSB-first stream ----
__device__ __forceinline__ uint32_t lit_token(uint32_t b, int* nb) {
  if (b < 144u) {
    *nb = 8;
    return __brev(0x30u + b) >> 24;
  }
  *nb = 9;
  return __brev(0x190u + b - 144u) >> 23;
}

__device__ __forceinline__ int match_bits(int run) {  // 3 <= run <= 128, distance 1
  const int l = run - 3;
  if (l < 8) return 7 + 5;
  const int e = 29 - __clz(l);  // floor(log2 l) - 2
  const int sym = 261 + 4 * e + ((l >> e) & 3);
  return (sym < 280 ? 7 : 8) + e + 5;
}

__device__ __forceinline__ uint32_t match_token(int run, int* nb) {
  const int l = run - 3;
  int e = 0, sym = 257 + l;
  if (l >= 8) {
    e = 29 - __clz(l);
    sym = 261 + 4 * e + ((l >> e) & 3);
  }
  uint32_t code;
  int cb;
  if (sym < 280) {
    cb = 7;
    code = __brev(static_cast<uint32_t>(sym - 256)) >> 25;
  } else {
    cb = 8;
    code = __brev(static_cast<uint32_t>(0xC0 + sym - 280)) >> 24;
  }
  *nb = cb + e + 5;
  return code | (static_cast<uint32_t>(l & ((1 << e) - 1)) << cb);
}

// ---- one chunk: filtered bytes in registers ----
struct Chunk {
  uint32_t w[32];  // Up-filtered bytes, zero beyond n
  int n;           // valid bytes
  int prev;        // filtered left neighbour, -1 at the start of a row
};

__device__ __forceinline__ void load_chunk(const uint8_t* __restrict__ img, int W, int r, int x0, bool vec, Chunk& c) {
  const uint8_t* cur = img + static_cast<size_t>(r) * W + x0;
  const uint8_t* up = cur - W;
  c.n = min(kPngChunk, W - x0);
  if (vec && c.n == kPngChunk) {
    const uint4* c4 = reinterpret_cast<const uint4*>(cur);
    uint4 a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = __ldg(c4 + k);
    if (r > 0) {
      const uint4* u4 = reinterpret_cast<const uint4*>(up);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint4 u = __ldg(u4 + k);
        a[k].x = __vsub4(a[k].x, u.x);
        a[k].y = __vsub4(a[k].y, u.y);
        a[k].z = __vsub4(a[k].z, u.z);
        a[k].w = __vsub4(a[k].w, u.w);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      c.w[4 * k + 0] = a[k].x;
      c.w[4 * k + 1] = a[k].y;
      c.w[4 * k + 2] = a[k].z;
      c.w[4 * k + 3] = a[k].w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      uint32_t word = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = 4 * k + q;
        if (j < c.n) {
          const uint32_t v = cur[j] - (r > 0 ? up[j] : 0u);
          word |= (v & 0xFFu) << (8 * q);
        }
      }
      c.w[k] = word;
    }
  }
  c.prev = -1;
  if (x0 > 0) c.prev = static_cast<int>((cur[-1] - (r > 0 ? up[-1] : 0u)) & 0xFFu);
}

// E3: bytes that belong to a run of >= 3 bytes equal to their left neighbour (they are covered by matches);
// G: bytes >= 144 (9-bit literals).  Both limited to the chunk's n bytes.
__device__ __forceinline__ void chunk_masks(const Chunk& c, u128* E3, u128* G) {
  uint32_t e[4] = {0, 0, 0, 0}, g[4] = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    // SWAR byte tests (the __vcmp*4 intrinsics are emulated with more instructions on sm_100):
    //   byte == 0 of x:  ~(((x & 0x7F..) + 0x7F..) | x) & 0x80..      byte >= 0x90 of w:  ((w & 0x7F..) + 0x70..) & w & 0x80..
    const uint32_t left = (c.w[k] << 8) | (k == 0 ? static_cast<uint32_t>(c.prev & 0xFF) : (c.w[k - 1] >> 24));
    const uint32_t x = c.w[k] ^ left;
    const uint32_t eq = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
    const uint32_t ge = ((c.w[k] & 0x7F7F7F7Fu) + 0x70707070u) & c.w[k] & 0x80808080u;
    e[k >> 3] |= ((((eq >> 7) * 0x01020408u) >> 24) & 0xFu) << (4 * (k & 7));
    g[k >> 3] |= ((((ge >> 7) * 0x01020408u) >> 24) & 0xFu) << (4 * (k & 7));
  }
  u128 E = make128(e[0], e[1], e[2], e[3]);
  if (c.prev < 0) E.lo &= ~1ull;
  const u128 nmask = low_mask(c.n);
  E = E & nmask;
  const u128 a = E & shr(E, 1) & shr(E, 2);
  *E3 = a | shl_small(a, 1) | shl_small(a, 2);
  *G = make128(g[0], g[1], g[2], g[3]) & nmask;
}

__device__ __forceinline__ uint32_t chunk_token_bits(const Chunk& c, u128 E3, u128 G) {
  const u128 lit = ~E3 & low_mask(c.n);
  uint32_t bits = 8u * popc128(lit) + popc128(lit & G);
  u128 m = E3;
  while (!is_zero(m)) {
    m = shr(m, ctz128(m));
    const u128 t = ~m;
    if (is_zero(t)) {
      bits += match_bits(128);
      break;
    }
    const int run = ctz128(t);
    bits += match_bits(run);
    m = shr(m, run);
  }
  return bits;
}

__device__ __forceinline__ uint32_t block_sum_u32(uint32_t v, uint32_t* s_tmp) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane_id() == 0) s_tmp[threadIdx.x >> 5] = v;
  __syncthreads();
  uint32_t t = 0;
#pragma unroll
  for (int k = 0; k < kPngThreads / 32; ++k) t += s_tmp[k];
  return t;
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPngThreads) k_png_count(const uint8_t* __restrict__ labels, PngGeom g, int vec,
                                                           PngWorkspace ws) {
  __shared__ uint32_t s_tmp[kPngThreads / 32];
  __shared__ unsigned long long s_tmp64[kPngThreads / 32];
  const int seg = blockIdx.x;
  const int img = seg / g.S, s = seg - img * g.S;
  const int r0 = s * g.R, rows = min(g.R, g.H - r0);
  const int L = rows * (g.W + 1);
  const int rr = threadIdx.x / g.cpr, ck = threadIdx.x - rr * g.cpr;
  uint32_t bits = 0, s1 = 0;
  unsigned long long s2 = 0;
  if (rr < rows) {
    Chunk c;
    const int x0 = ck * kPngChunk;
    load_chunk(labels + static_cast<size_t>(img) * g.H * g.W, g.W, r0 + rr, x0, vec != 0, c);
    u128 E3, G;
    chunk_masks(c, &E3, &G);
    bits = chunk_token_bits(c, E3, G);
    uint32_t sb = 0, sj = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      sb = __dp4a(c.w[k], 0x01010101u, sb);
      const uint32_t j = 4u * k;
      sj = __dp4a(c.w[k], j | ((j + 1) << 8) | ((j + 2) << 16) | ((j + 3) << 24), sj);
    }
    const int wt0 = L - (rr * (g.W + 1) + 1 + x0);  // Adler weight of the chunk's first byte
    s1 = sb;
    s2 = static_cast<unsigned long long>(wt0) * sb - sj;
    if (ck == 0) {  // the row's filter byte
      bits += 8;
      s1 += kFilterUp;
      s2 += static_cast<unsigned long long>(kFilterUp) * (L - rr * (g.W + 1));
    }
  }
  ws.chunk_bits[static_cast<size_t>(seg) * kPngThreads + threadIdx.x] = static_cast<uint16_t>(bits);
  const uint32_t tb = block_sum_u32(bits, s_tmp);
  const uint32_t t1 = block_sum_u32(s1, s_tmp);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  __syncthreads();
  if (lane_id() == 0) s_tmp64[threadIdx.x >> 5] = s2;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t2 = 0;
    for (int k = 0; k < kPngThreads / 32; ++k) t2 += s_tmp64[k];
    ws.seg_bits[seg] = tb;
    ws.seg_s1[seg] = t1;
    ws.seg_s2[seg] = t2;
  }
}

__device__ __forceinline__ void put_be32(uint8_t* p, uint32_t v) {
  p[0] = static_cast<uint8_t>(v >> 24);
  p[1] = static_cast<uint8_t>(v >> 16);
  p[2] = static_cast<uint8_t>(v >> 8);
  p[3] = static_cast<uint8_t>(v);
}

__device__ uint32_t crc32_bitwise(const uint8_t* p, int n) {
  uint32_t c = 0xFFFFFFFFu;
  for (int i = 0; i < n; ++i) {
    c ^= p[i];
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? kCrcPoly : 0u);
  }
  return c ^ 0xFFFFFFFFu;
}

constexpr int kOffsetsThreads = 1024;  // one CTA; a warp per image, 32 images in flight

__global__ void __launch_bounds__(kOffsetsThreads) k_png_offsets(PngGeom g, int n_images, PngWorkspace ws, uint8_t* out,
                                                             unsigned long long capacity, long long* offsets) {
  // one warp per image, lanes over its segments: sizes and modes, chunk offsets by a warp scan, the Adler-32 state of every
  // segment start from the prefix sums of the byte sums (A_s = 1 + sum_{k<s} s1_k; B = sum_s (L_s * A_s + s2_s), all mod 65521)
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  for (int img = warp; img < n_images; img += kOffsetsThreads / 32) {
    uint32_t pos = 8 + 25;
    unsigned long long a_run = 1, b_acc = 0;
    for (int s0 = 0; s0 < g.S; s0 += 32) {
      const int s = s0 + lane;
      const bool in = s < g.S;
      const int seg = img * g.S + (in ? s : 0);
      const int rows = min(g.R, g.H - (in ? s : 0) * g.R);
      const uint32_t L = static_cast<uint32_t>(rows) * (g.W + 1);
      const uint32_t fixed = (ws.seg_bits[seg] + 13 + 7) / 8 + 4;
      const uint32_t stored = 5 + L;
      const uint32_t mode = fixed > stored ? 1u : 0u;
      const uint32_t d = (s == 0 ? 2u : 0u) + (mode ? stored : fixed) + (s == g.S - 1 ? 6u : 0u);
      const uint32_t bytes = in ? 12 + d : 0u;
      unsigned long long s1 = in ? ws.seg_s1[seg] : 0ull;
      const unsigned long long s2 = in ? ws.seg_s2[seg] % kAdlerMod : 0ull;
      uint32_t incl = bytes;
      unsigned long long s1_incl = s1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        const unsigned long long u = __shfl_up_sync(0xffffffffu, s1_incl, o);
        if (lane >= o) {
          incl += t;
          s1_incl += u;
        }
      }
      if (in) {
        ws.seg_off[seg] = pos + incl - bytes;
        ws.seg_len[seg] = d | (mode << 31);
      }
      const unsigned long long a_s = (a_run + (s1_incl - s1)) % kAdlerMod;   // Adler A at the start of segment s
      unsigned long long term = in ? ((L % kAdlerMod) * a_s + s2) % kAdlerMod : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
      b_acc = (b_acc + term) % kAdlerMod;
      pos += __shfl_sync(0xffffffffu, incl, 31);
      a_run = (a_run + __shfl_sync(0xffffffffu, s1_incl, 31)) % kAdlerMod;
    }
    if (lane == 0) {
      ws.adler[img] = static_cast<uint32_t>((b_acc << 16) | a_run);
      offsets[img + 1] = pos + 12;
    }
  }
  __syncthreads();
  {  // inclusive scan of the file sizes over the images, 1024 at a time
    __shared__ long long s_scan[kOffsetsThreads / 32];
    __shared__ long long s_carry;
    if (threadIdx.x == 0) {
      s_carry = 0;
      offsets[0] = 0;
    }
    __syncthreads();
    for (int base = 0; base < n_images; base += kOffsetsThreads) {
      const int i = base + threadIdx.x;
      long long v = i < n_images ? offsets[i + 1] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      if (lane == 31) s_scan[warp] = v;
      __syncthreads();
      long long add = s_carry;
      for (int k = 0; k < warp; ++k) add += s_scan[k];
      v += add;
      if (i < n_images) offsets[i + 1] = v;
      __syncthreads();
      if (threadIdx.x == kOffsetsThreads - 1) s_carry = v;
      __syncthreads();
    }
  }
  if (static_cast<unsigned long long>(offsets[n_images]) > capacity) return;
  for (int img = threadIdx.x; img < n_images; img += blockDim.x) {
    uint8_t* f = out + offsets[img];
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    for (int k = 0; k < 8; ++k) f[k] = sig[k];
    uint8_t hdr[17] = {'I', 'H', 'D', 'R', 0, 0, 0, 0, 0, 0, 0, 0, 8, 0, 0, 0, 0};
    put_be32(hdr + 4, static_cast<uint32_t>(g.W));
    put_be32(hdr + 8, static_cast<uint32_t>(g.H));
    put_be32(f + 8, 13);
    for (int k = 0; k < 17; ++k) f[12 + k] = hdr[k];
    put_be32(f + 29, crc32_bitwise(hdr, 17));
    uint8_t* e = out + offsets[img + 1] - 12;
    const uint8_t iend[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xAE, 0x42, 0x60, 0x82};
    for (int k = 0; k < 12; ++k) e[k] = iend[k];
  }
}

// Bit writer of one thread into the CTA's zero-initialised shared bit buffer.  The first and the last word a thread
// touches may be shared with its neighbours (atomicOr); the words in between are its own.
struct BitWriter {
  uint32_t* buf;
  unsigned long long acc;
  int fill;
  uint32_t word;
  bool first;
  __device__ __forceinline__ void init(uint32_t* b, uint32_t bitpos) {
    buf = b;
    word = bitpos >> 5;
    fill = static_cast<int>(bitpos & 31u);
    acc = 0;
    first = true;
  }
  __device__ __forceinline__ void put(uint32_t v, int nb) {
    acc |= static_cast<unsigned long long>(v) << fill;
    fill += nb;
    if (fill >= 32) {
      if (first) {
        atomicOr(buf + word, static_cast<uint32_t>(acc));
        first = false;
      } else {
        buf[word] = static_cast<uint32_t>(acc);
      }
      ++word;
      acc >>= 32;
      fill -= 32;
    }
  }
  __device__ __forceinline__ void finish() {
    if (fill > 0 && static_cast<uint32_t>(acc) != 0u) atomicOr(buf + word, static_cast<uint32_t>(acc));
  }
};

// VAR 0: per-lane token loop with an inner literal loop (lanes diverge); VAR 1 (default): one token per iteration, branch-free body.
template <int VAR>
__global__ void __launch_bounds__(kPngThreads, 4) k_png_emit(const uint8_t* __restrict__ labels, PngGeom g, int vec,
                                                          PngWorkspace ws, uint8_t* __restrict__ out,
                                                          unsigned long long capacity,
                                                          const long long* __restrict__ offsets, int n_images) {
  extern __shared__ __align__(16) uint32_t s_buf[];
  __shared__ uint32_t s_tab[256];
  __shared__ uint32_t s_crc[kPngThreads];
  __shared__ uint32_t s_warp[kPngThreads / 32];
  if (static_cast<unsigned long long>(offsets[n_images]) > capacity) return;
  const int seg = blockIdx.x;
  const int img = seg / g.S, s = seg - img * g.S;
  const int r0 = s * g.R, rows = min(g.R, g.H - r0);
  const uint32_t L = static_cast<uint32_t>(rows) * (g.W + 1);
  const uint32_t d = ws.seg_len[seg] & 0x7FFFFFFFu;
  const bool stored = (ws.seg_len[seg] >> 31) != 0;
  const unsigned long long dst = static_cast<unsigned long long>(offsets[img]) + ws.seg_off[seg];
  const uint32_t pad = static_cast<uint32_t>(dst & 3ull);
  const uint32_t total = pad + 12 + d;  // bytes of the staging buffer in use
  const uint32_t n_words = (total + 3) / 4;
  uint8_t* s_bytes = reinterpret_cast<uint8_t*>(s_buf);
  const uint32_t zl = s == 0 ? 2u : 0u;
  const uint32_t data0 = pad + 8 + zl;  // first byte of the DEFLATE data

  for (uint32_t k = threadIdx.x; k < n_words; k += kPngThreads) s_buf[k] = 0;
  {  // CRC-32 byte table
    uint32_t c = threadIdx.x;
#pragma unroll
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? kCrcPoly : 0u);
    s_tab[threadIdx.x] = c;
  }

  const int rr = threadIdx.x / g.cpr, ck = threadIdx.x - rr * g.cpr;
  const bool active = rr < rows;
  const int x0 = ck * kPngChunk;
  const uint8_t* image = labels + static_cast<size_t>(img) * g.H * g.W;
  Chunk c;
  c.n = 0;
  if (active && !stored) load_chunk(image, g.W, r0 + rr, x0, vec != 0, c);

  // exclusive scan of the chunks' bit counts
  const uint32_t my_bits = ws.chunk_bits[static_cast<size_t>(seg) * kPngThreads + threadIdx.x];
  uint32_t incl = my_bits;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane_id() >= o) incl += t;
  }
  if (lane_id() == 31) s_warp[threadIdx.x >> 5] = incl;
  __syncthreads();  // also: zero fill and CRC table complete
  uint32_t base = 0;
  for (int k = 0; k < (threadIdx.x >> 5); ++k) base += s_warp[k];
  const uint32_t bit0 = data0 * 8 + 3 + base + incl - my_bits;

  if (!stored) {
    if (threadIdx.x == 0) atomicOr(s_buf + ((data0 * 8 + 1) >> 5), 1u << ((data0 * 8 + 1) & 31u));  // BTYPE = 01
    if (active) {
      BitWriter bw;
      bw.init(s_buf, bit0);
      int nb;
      if (ck == 0) {
        const uint32_t v = lit_token(kFilterUp, &nb);
        bw.put(v, nb);
      }
      u128 E3, G;
      chunk_masks(c, &E3, &G);
      const uint8_t* cur = image + static_cast<size_t>(r0 + rr) * g.W + x0;
      const bool has_up = r0 + rr > 0;
      int i = 0;
      if (VAR == 0) {
        while (i < c.n) {
          const u128 t = shr(E3, i);
          if (t.lo & 1ull) {
            const u128 z = ~t;
            const int run = is_zero(z) ? 128 : ctz128(z);
            const uint32_t v = match_token(run, &nb);
            bw.put(v, nb);
            i += run;
          } else {
            int nl = is_zero(t) ? 128 : ctz128(t);
            nl = min(nl, c.n - i);
            for (int j = 0; j < nl; ++j) {
              const uint32_t b = (__ldg(cur + i + j) - (has_up ? __ldg(cur + i + j - g.W) : 0u)) & 0xFFu;
              const uint32_t v = lit_token(b, &nb);
              bw.put(v, nb);
            }
            i += nl;
          }
        }
      } else {
        // One token per iteration, both kinds evaluated and selected: the lanes of a warp (32 different chunks) stay
        // converged, the warp runs max-over-lanes iterations instead of the union of 32 different branch sequences.
        while (i < c.n) {
          const u128 t = shr(E3, i);
          const bool is_match = (t.lo & 1ull) != 0ull;
          const u128 z = ~t;
          const int run = is_zero(z) ? 128 : ctz128(z);
          uint32_t b = 0;
          if (!is_match) b = (__ldg(cur + i) - (has_up ? __ldg(cur + i - g.W) : 0u)) & 0xFFu;
          int nbm, nbl;
          const uint32_t vm = match_token(max(run, 3), &nbm);
          const uint32_t vl = lit_token(b, &nbl);
          bw.put(is_match ? vm : vl, is_match ? nbm : nbl);
          i += is_match ? run : 1;
        }
      }
      bw.finish();
    }
  } else {
    // Stored segment: the raw stream (filter byte + Up-filtered row, row after row) is assembled one aligned shared word per
    // thread from global memory.  (Writing each thread's own 128 chunk bytes put all lanes of a warp on one bank -- chunks
    // are 128 bytes apart -- a 32-way conflict on every byte store: 13 % of the kernel on maps with a few stored segments.)
    const uint32_t raw_base = data0 + 5;
    const int Wp = g.W + 1;
    const uint8_t* seg_rows = image + static_cast<size_t>(r0) * g.W;
    for (uint32_t wi = raw_base / 4 + threadIdx.x; wi < (raw_base + L + 3) / 4; wi += kPngThreads) {
      int o = static_cast<int>(wi * 4) - static_cast<int>(raw_base);
      int r = o >= 0 ? o / Wp : 0;
      int col = o >= 0 ? o - r * Wp : o;  // negative: bytes in front of the raw stream (header, written later)
      uint32_t word = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (col >= 0 && r < rows) {
          uint32_t v = kFilterUp;
          if (col > 0) {
            const uint8_t* p = seg_rows + static_cast<size_t>(r) * g.W + (col - 1);
            v = (__ldg(p) - ((r0 + r > 0) ? __ldg(p - g.W) : 0u)) & 0xFFu;
          }
          word |= v << (8 * b);
        }
        if (++col == Wp) {
          col = 0;
          ++r;
        }
      }
      s_buf[wi] = word;
    }
  }
  __syncthreads();

  if (threadIdx.x == 0) {  // framing bytes
    put_be32(s_bytes + pad, d);
    s_bytes[pad + 4] = 'I';
    s_bytes[pad + 5] = 'D';
    s_bytes[pad + 6] = 'A';
    s_bytes[pad + 7] = 'T';
    if (s == 0) {
      s_bytes[pad + 8] = 0x78;
      s_bytes[pad + 9] = 0x01;
    }
    uint32_t end = pad + 8 + d;  // one past the data
    if (s == g.S - 1) {
      end -= 6;
      s_bytes[end] = 0x03;
      s_bytes[end + 1] = 0x00;
      put_be32(s_bytes + end + 2, ws.adler[img]);
    }
    if (stored) {
      s_bytes[data0] = 0x00;
      s_bytes[data0 + 1] = static_cast<uint8_t>(L);
      s_bytes[data0 + 2] = static_cast<uint8_t>(L >> 8);
      s_bytes[data0 + 3] = static_cast<uint8_t>(~L);
      s_bytes[data0 + 4] = static_cast<uint8_t>((~L) >> 8);
    } else {  // sync flush 00 00 FF FF: the zeros are there already
      s_bytes[end - 2] = 0xFF;
      s_bytes[end - 1] = 0xFF;
    }
  }
  __syncthreads();

  // CRC-32 of type + data: right-aligned equal pieces, finalised piece CRCs, shift-and-xor tree (zlib's crc32_combine)
  const uint32_t n_crc = 4 + d;
  const uint8_t* crc_src = s_bytes + pad + 4;
  int lv = 0;
  while (lv < kCrcLevels && (64u << lv) < n_crc) ++lv;
  const int tn = 1 << lv;
  const uint32_t m = (n_crc + tn - 1) / tn;  // < kCrcMaxPiece
  if (threadIdx.x < tn) {
    const int hi = static_cast<int>(n_crc) - (tn - 1 - static_cast<int>(threadIdx.x)) * static_cast<int>(m);
    const int lo = max(hi - static_cast<int>(m), 0);
    uint32_t cr = 0xFFFFFFFFu;
    for (int k = lo; k < hi; ++k) cr = s_tab[(cr ^ crc_src[k]) & 0xFFu] ^ (cr >> 8);
    s_crc[threadIdx.x] = hi > lo ? cr ^ 0xFFFFFFFFu : 0u;
  }
  __syncthreads();
  for (int l = 0; l < lv; ++l) {  // thread j combines pieces (2j, 2j + 1) << l: the active threads stay contiguous
    const int left = static_cast<int>(threadIdx.x) << (l + 1);
    if (left < tn) s_crc[left] = multmodp(c_crc_shift.v[l][m], s_crc[left]) ^ s_crc[left + (1 << l)];
    __syncthreads();
  }
  if (threadIdx.x == 0) put_be32(s_bytes + pad + 8 + d, s_crc[0]);
  __syncthreads();

  // copy: shared word k <-> destination word (dst - pad) / 4 + k; the partial end words byte by byte
  uint8_t* gbase = out + (dst - pad);
  uint32_t* gw = reinterpret_cast<uint32_t*>(gbase);
  const uint32_t first_full = pad ? 1u : 0u;
  const uint32_t last_full = total / 4;  // words [first_full, last_full) are complete
  for (uint32_t k = first_full + threadIdx.x; k < last_full; k += kPngThreads) gw[k] = s_buf[k];
  if (threadIdx.x == 0) {
    if (pad)
      for (uint32_t b = pad; b < 4 && b < total; ++b) gbase[b] = s_bytes[b];
    for (uint32_t b = max(last_full * 4, pad ? 4u : 0u); b < total; ++b) gbase[b] = s_bytes[b];
  }
}

}  // namespace
}  // namespace hiast

using namespace hiast;

static int g_png_variant = 1;
extern "C" int hiast_debug_png_variant(int v) {
  g_png_variant = v;
  return HIAST_OK;
}

extern "C" size_t hiast_png_workspace_bytes(int n_images, int H, int W) {
  PngGeom g;
  if (n_images < 0 || !png_geom(H, W, &g)) return 0;
  return png_carve(nullptr, n_images, g, nullptr);
}

extern "C" size_t hiast_png_max_bytes(int H, int W) {
  PngGeom g;
  if (!png_geom(H, W, &g)) return 0;
  return static_cast<size_t>(8 + 25 + 12) + static_cast<size_t>(g.S) * (12 + 5) + static_cast<size_t>(H) * (W + 1) + 2 + 6;
}

extern "C" int hiast_png_segments(int H, int W) {
  PngGeom g;
  if (!png_geom(H, W, &g)) return 0;
  return g.S;
}

extern "C" int hiast_png_encode(const uint8_t* labels, int n_images, int H, int W, uint8_t* out, size_t out_capacity,
                                int64_t* offsets, void* workspace, size_t workspace_bytes, void* stream) {
  PngGeom g;
  if (!labels || !out || !offsets || !workspace || n_images < 0) return HIAST_ERR_INVALID_ARG;
  if (!png_geom(H, W, &g)) return HIAST_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(out) % 4 != 0 || reinterpret_cast<uintptr_t>(workspace) % 256 != 0)
    return HIAST_ERR_INVALID_ARG;
  if (static_cast<long long>(n_images) * g.S > 0x7FFFFFFFll) return HIAST_ERR_UNSUPPORTED;
  PngWorkspace ws;
  if (png_carve(workspace, n_images, g, &ws) > workspace_bytes) return HIAST_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (n_images == 0) {
    HIAST_CUDA_TRY(cudaMemsetAsync(offsets, 0, sizeof(int64_t), st));
    return HIAST_OK;
  }
  const int vec = (W % 16 == 0) && (reinterpret_cast<uintptr_t>(labels) % 16 == 0);
  const int n_seg = n_images * g.S;
  k_png_count<<<n_seg, kPngThreads, 0, st>>>(labels, g, vec, ws);
  HIAST_CHECK_LAUNCH();
  k_png_offsets<<<1, kOffsetsThreads, 0, st>>>(g, n_images, ws, out, static_cast<unsigned long long>(out_capacity),
                                          reinterpret_cast<long long*>(offsets));
  HIAST_CHECK_LAUNCH();
  const size_t smem = (static_cast<size_t>(g.Lmax) + 5 + 3 + 12 + 2 + 6 + 15) / 16 * 16;
  if (smem + 3 * 1024 > 48 * 1024) return HIAST_ERR_UNSUPPORTED;  // cannot happen: Lmax <= 256 * 129
  if (g_png_variant == 0)
    k_png_emit<0><<<n_seg, kPngThreads, smem, st>>>(labels, g, vec, ws, out, static_cast<unsigned long long>(out_capacity),
                                                   reinterpret_cast<const long long*>(offsets), n_images);
  else
    k_png_emit<1><<<n_seg, kPngThreads, smem, st>>>(labels, g, vec, ws, out, static_cast<unsigned long long>(out_capacity),
                                                   reinterpret_cast<const long long*>(offsets), n_images);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}
