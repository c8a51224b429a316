// Memory-system diagnostics for IAS phase A (development tool, not part of the library).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/membench tools/membench.cu
//   ./tools/membench [images]
//
// Phase A reads 19 channel planes of 8 MiB each per image (NCHW) and is stuck near 5.1 TB/s of DRAM traffic
// with the issue slots only 48 % busy (packed-math kernel).  These kernels take the production pipeline
// (256 threads, every thread owns 19 x 16 B of shared memory, cp.async of the next tile while the current one
// is consumed) and switch its parts on and off to find which one holds the memory system back:
//   pattern  NCHW planes | tile-major linear addresses
//   work     trivial max | the production packed softmax / arg-max
//   writes   none | conf f32 + label u8
//   atomics  none | one global RED per pixel into a [2][19][4420] histogram
// plus a plain grid-stride 128-bit read kernel and a copy kernel as machine references.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../hiast_b200/csrc/api.cu"
#include "../hiast_b200/csrc/ias_phase_a.cu"

using namespace hiast;

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                      \
    }                                                                               \
  } while (0)

constexpr int kC = 19;

struct MArgs {
  const float* logits;
  float* conf;
  uint8_t* label;
  uint32_t* hist;
  int n_images;
  int64_t HW;
  int tiles_per_image;
  int n_tiles;
  int chunk;
  unsigned* sched;
  float* sink;
};

__global__ void k_init(float* p, size_t n, unsigned seed) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    unsigned h = static_cast<unsigned>(i) * 2654435761u ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    // roughly N(0, 3^2): sum of 4 uniforms with 24-bit resolution (no exact ties between channels)
    unsigned h2 = h * 747796405u + 2891336453u;
    h2 ^= h2 >> 16; h2 *= 2246822519u; h2 ^= h2 >> 13;
    float u = ((h & 0xffff) + (h >> 16) + (h2 & 0xffff) + (h2 >> 16)) * (1.0f / 65535.0f) - 2.0f;
    p[i] = u * 5.2f + (h2 & 1023) * 1e-7f;
  }
}

// PATTERN 0 NCHW, 1 tile-major linear.  WORK 0 trivial, 1 packed softmax.  DEPTH: tiles of prefetch distance is 1.
template <int PATTERN, int WORK, int WRITES, int ATOM, int MINB>
__global__ void __launch_bounds__(256, MINB) k_stream(MArgs a) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  float4* s_stage = reinterpret_cast<float4*>(s_raw);
  __shared__ int s_slot[2];
  const int HW4 = static_cast<int>(a.HW / 4);
  float4* my = s_stage + threadIdx.x;
  const unsigned my_u32 = static_cast<unsigned>(__cvta_generic_to_shared(my));
  auto prefetch = [&](int t) {
    const int img = t / a.tiles_per_image, tile = t - img * a.tiles_per_image;
    const char* src;
    size_t stride;
    if (PATTERN == 0) {
      src = reinterpret_cast<const char*>(a.logits + static_cast<size_t>(img) * kC * a.HW) +
            (static_cast<size_t>(tile) * 256 + threadIdx.x) * 16;
      stride = static_cast<size_t>(a.HW) * 4;
    } else {
      src = reinterpret_cast<const char*>(a.logits) + (static_cast<size_t>(t) * kC * 256 + threadIdx.x) * 16;
      stride = 256 * 16;
    }
#pragma unroll
    for (int c = 0; c < kC; ++c) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(my_u32 + c * 256 * 16), "l"(src) : "memory");
      src += stride;
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  const int n_chunks = (a.n_tiles + a.chunk - 1) / a.chunk;
  int cur = blockIdx.x, nxt = blockIdx.x + gridDim.x, par = 0;
  if (cur < n_chunks) prefetch(cur * a.chunk);
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  float keep = 0.f;
  while (cur < n_chunks) {
    if (threadIdx.x == 0) s_slot[par] = static_cast<int>(atomicAdd(a.sched, 1u)) + 2 * static_cast<int>(gridDim.x);
    const int t0 = cur * a.chunk, t1 = min(t0 + a.chunk, a.n_tiles);
    for (int t = t0; t < t1; ++t) {
      int nt = t + 1;
      bool has_next = true;
      if (nt == t1) {
        has_next = nxt < n_chunks;
        nt = nxt * a.chunk;
      }
      float v[4][kC];
#pragma unroll
      for (int c = 0; c < kC; ++c) {
        const float4 q = my[c * 256];
        v[0][c] = q.x; v[1][c] = q.y; v[2][c] = q.z; v[3][c] = q.w;
      }
      float guard = 0.f;
#pragma unroll
      for (int c = 0; c < kC; ++c) guard = fmaxf(guard, v[0][c]);
      if (has_next && guard == guard) prefetch(nt);
      float cf[4];
      int lb[4];
      if (WORK == 1) {
        bool tie[4];
        softmax_argmax_pair<kC>(v[0], v[1], cf[0], cf[1], lb[0], lb[1], tie[0], tie[1]);
        softmax_argmax_pair<kC>(v[2], v[3], cf[2], cf[3], lb[2], lb[3], tie[2], tie[3]);
        if (tie[0] | tie[1] | tie[2] | tie[3]) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (tie[j]) softmax_argmax<kC>(v[j], cf[j], lb[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float m = v[j][0];
#pragma unroll
          for (int c = 1; c < kC; ++c) m = fmaxf(m, v[j][c]);
          cf[j] = 1.0f / (1.0f + fabsf(m));
          lb[j] = __float_as_int(m) & 15;
        }
      }
      const int img = t / a.tiles_per_image, tile = t - img * a.tiles_per_image;
      const size_t o4 = static_cast<size_t>(img) * HW4 + tile * 256 + threadIdx.x;
      if (WRITES) {
        reinterpret_cast<float4*>(a.conf)[o4] = make_float4(cf[0], cf[1], cf[2], cf[3]);
        reinterpret_cast<uchar4*>(a.label)[o4] = make_uchar4(lb[0], lb[1], lb[2], lb[3]);
      } else {
        keep += cf[0] + cf[1] + cf[2] + cf[3] + lb[0] + lb[3];
      }
      if (ATOM) {
        uint32_t* g = a.hist + static_cast<size_t>(img & 1) * kC * 4420;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int bin = min(max(static_cast<int>(fp16_key(cf[j])) - 0x2ABD, 0), 4419);
          atomicAdd(g + lb[j] * 4420 + bin, 1u);
        }
      }
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();
    const int nn = s_slot[par];
    par ^= 1;
    cur = nxt;
    nxt = nn;
  }
  if (!WRITES && keep == 123.456f) a.sink[0] = keep;
}

// Two tiles in flight per thread: 2 x 19 x 16 B of shared memory per thread, one CTA of 256 threads per ~152 KB.
// (Checks whether more bytes in flight per SM buy bandwidth: 1 CTA x 2 stages = the same 152 KB as 2 CTAs x 1.)
template <int WORK>
__global__ void __launch_bounds__(512, 1) k_stream_wide(MArgs a) {
  // 512 threads x 1 stage: 16 warps in ONE CTA (same bytes in flight as 2 x 256), to see the effect of CTA shape
  extern __shared__ __align__(16) unsigned char s_raw[];
  float4* s_stage = reinterpret_cast<float4*>(s_raw);
  __shared__ int s_slot[2];
  const int HW4 = static_cast<int>(a.HW / 4);
  float4* my = s_stage + threadIdx.x;
  const unsigned my_u32 = static_cast<unsigned>(__cvta_generic_to_shared(my));
  const int tpi = a.tiles_per_image / 2, n_tiles = a.n_tiles / 2;   // tiles of 2048 pixels
  auto prefetch = [&](int t) {
    const int img = t / tpi, tile = t - img * tpi;
    const char* src = reinterpret_cast<const char*>(a.logits + static_cast<size_t>(img) * kC * a.HW) +
                      (static_cast<size_t>(tile) * 512 + threadIdx.x) * 16;
    const size_t stride = static_cast<size_t>(a.HW) * 4;
#pragma unroll
    for (int c = 0; c < kC; ++c) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(my_u32 + c * 512 * 16), "l"(src) : "memory");
      src += stride;
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  const int n_chunks = (n_tiles + a.chunk - 1) / a.chunk;
  int cur = blockIdx.x, nxt = blockIdx.x + gridDim.x, par = 0;
  if (cur < n_chunks) prefetch(cur * a.chunk);
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  float keep = 0.f;
  while (cur < n_chunks) {
    if (threadIdx.x == 0) s_slot[par] = static_cast<int>(atomicAdd(a.sched, 1u)) + 2 * static_cast<int>(gridDim.x);
    const int t0 = cur * a.chunk, t1 = min(t0 + a.chunk, n_tiles);
    for (int t = t0; t < t1; ++t) {
      int nt = t + 1;
      bool has_next = true;
      if (nt == t1) {
        has_next = nxt < n_chunks;
        nt = nxt * a.chunk;
      }
      float v[4][kC];
#pragma unroll
      for (int c = 0; c < kC; ++c) {
        const float4 q = my[c * 512];
        v[0][c] = q.x; v[1][c] = q.y; v[2][c] = q.z; v[3][c] = q.w;
      }
      float guard = 0.f;
#pragma unroll
      for (int c = 0; c < kC; ++c) guard = fmaxf(guard, v[0][c]);
      if (has_next && guard == guard) prefetch(nt);
      float cf[4];
      int lb[4];
      if (WORK == 1) {
        bool tie[4];
        softmax_argmax_pair<kC>(v[0], v[1], cf[0], cf[1], lb[0], lb[1], tie[0], tie[1]);
        softmax_argmax_pair<kC>(v[2], v[3], cf[2], cf[3], lb[2], lb[3], tie[2], tie[3]);
        if (tie[0] | tie[1] | tie[2] | tie[3]) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (tie[j]) softmax_argmax<kC>(v[j], cf[j], lb[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float m = v[j][0];
#pragma unroll
          for (int c = 1; c < kC; ++c) m = fmaxf(m, v[j][c]);
          cf[j] = m;
          lb[j] = 0;
        }
      }
      keep += cf[0] + cf[1] + cf[2] + cf[3] + lb[0] + lb[3];
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();
    const int nn = s_slot[par];
    par ^= 1;
    cur = nxt;
    nxt = nn;
  }
  if (keep == 123.456f) a.sink[0] = keep + HW4;
}

// machine references
__global__ void __launch_bounds__(256) k_read(const float4* __restrict__ p, size_t n4, float* sink) {
  float acc = 0.f;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  for (; i + 7 * stride < n4; i += 8 * stride) {
    float4 q[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) q[k] = __ldcs(p + i + k * stride);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += q[k].x + q[k].y + q[k].z + q[k].w;
  }
  for (; i < n4; i += stride) {
    const float4 q = __ldcs(p + i);
    acc += q.x + q.y + q.z + q.w;
  }
  if (acc == 123.456f) sink[0] = acc;
}

// NCHW plane pattern with plain LDG at high occupancy: thread reads 19 x 16 B (one per plane), sums.
__global__ void __launch_bounds__(256) k_read_planes(const float* __restrict__ logits, int64_t HW, int n_images, float* sink) {
  const int HW4 = static_cast<int>(HW / 4);
  const long long total = static_cast<long long>(n_images) * HW4;
  float acc = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int img = static_cast<int>(i / HW4), p4 = static_cast<int>(i - static_cast<long long>(img) * HW4);
    const float4* src = reinterpret_cast<const float4*>(logits + static_cast<size_t>(img) * kC * HW) + p4;
    float4 q[kC];
#pragma unroll
    for (int c = 0; c < kC; ++c) q[c] = __ldcs(src + static_cast<size_t>(c) * HW4);
#pragma unroll
    for (int c = 0; c < kC; ++c) acc += q[c].x + q[c].y + q[c].z + q[c].w;
  }
  if (acc == 123.456f) sink[0] = acc;
}

__global__ void __launch_bounds__(256) k_copy(const float4* __restrict__ p, float4* __restrict__ o, size_t n4) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = __ldcs(p + i + k * stride);
#pragma unroll
    for (int k = 0; k < 4; ++k) __stcs(o + i + k * stride, q[k]);
  }
  for (; i < n4; i += stride) o[i] = p[i];
}

template <typename F>
float time_ms(F fn, int iters = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  fn();
  fn();
  CK(cudaDeviceSynchronize());
  std::vector<float> ts;
  for (int i = 0; i < iters; ++i) {
    CK(cudaEventRecord(e0));
    fn();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ts.push_back(ms);
  }
  std::sort(ts.begin(), ts.end());
  return ts[ts.size() / 2];
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 32;
  const int64_t HW = 1024 * 2048;
  const size_t elems = static_cast<size_t>(n) * kC * HW;
  float *logits, *conf, *sink, *copy_dst;
  uint8_t* label;
  uint32_t* hist;
  unsigned* sched;
  CK(cudaMalloc(&logits, elems * 4));
  CK(cudaMalloc(&copy_dst, elems * 4));
  CK(cudaMalloc(&conf, static_cast<size_t>(n) * HW * 4));
  CK(cudaMalloc(&label, static_cast<size_t>(n) * HW));
  CK(cudaMalloc(&hist, 2 * kC * 4420 * 4));
  CK(cudaMalloc(&sched, 4));
  CK(cudaMalloc(&sink, 4));
  k_init<<<148 * 8, 256>>>(logits, elems, 12345u);
  CK(cudaMemset(hist, 0, 2 * kC * 4420 * 4));
  CK(cudaDeviceSynchronize());
  const double bytes_in = static_cast<double>(elems) * 4;
  auto report = [&](const char* name, float ms, double bytes) {
    printf("%-58s %8.3f ms  %7.1f GB/s  (%.1f us/img)\n", name, ms, bytes / ms / 1e6, ms * 1e3 / n);
    fflush(stdout);
  };
  // machine references
  for (int occ : {2, 4, 8}) {
    char nm[128];
    snprintf(nm, sizeof nm, "read linear LDG.128 x8 unroll, %d CTAs/SM", occ);
    report(nm, time_ms([&] { k_read<<<148 * occ, 256>>>(reinterpret_cast<const float4*>(logits), elems / 4, sink); }), bytes_in);
  }
  for (int occ : {2, 4, 8}) {
    char nm[128];
    snprintf(nm, sizeof nm, "read NCHW planes LDG.128 x19, %d CTAs/SM", occ);
    report(nm, time_ms([&] { k_read_planes<<<148 * occ, 256>>>(logits, HW, n, sink); }), bytes_in);
  }
  report("copy LDG/STG.128 (read+write bytes)", time_ms([&] { k_copy<<<148 * 8, 256>>>(reinterpret_cast<const float4*>(logits), reinterpret_cast<float4*>(copy_dst), elems / 4); }), 2 * bytes_in);
  CK(cudaFree(copy_dst));

  MArgs a;
  a.logits = logits; a.conf = conf; a.label = label; a.hist = hist; a.n_images = n; a.HW = HW;
  a.tiles_per_image = static_cast<int>(HW / 1024);
  a.n_tiles = a.tiles_per_image * n;
  a.sched = sched; a.sink = sink;
  const size_t smem = 16 * kC * 256;
  auto run = [&](auto kern, const char* name, int chunk, int grid, size_t sm, double bytes, int threads = 256) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm)));
    a.chunk = chunk;
    const float ms = time_ms([&] {
      CK(cudaMemsetAsync(sched, 0, 4));
      kern<<<grid, threads, sm>>>(a);
    });
    char nm[160];
    snprintf(nm, sizeof nm, "%s chunk=%d grid=%d", name, chunk, grid);
    report(nm, ms, bytes);
  };
  const double bytes_w = bytes_in + static_cast<double>(n) * HW * 5;
  run(k_stream<0, 0, 0, 0, 2>, "sp NCHW   trivial  nowrite noatom", 8, 296, smem, bytes_in);
  run(k_stream<1, 0, 0, 0, 2>, "sp linear trivial  nowrite noatom", 8, 296, smem, bytes_in);
  run(k_stream<0, 1, 0, 0, 2>, "sp NCHW   softmax  nowrite noatom", 8, 296, smem, bytes_in);
  run(k_stream<1, 1, 0, 0, 2>, "sp linear softmax  nowrite noatom", 8, 296, smem, bytes_in);
  run(k_stream<0, 1, 1, 0, 2>, "sp NCHW   softmax  write   noatom", 8, 296, smem, bytes_w);
  run(k_stream<0, 1, 1, 1, 2>, "sp NCHW   softmax  write   atom  ", 8, 296, smem, bytes_w);
  run(k_stream<0, 0, 1, 0, 2>, "sp NCHW   trivial  write   noatom", 8, 296, smem, bytes_w);
  for (int chunk : {1, 2, 4, 16, 64}) run(k_stream<0, 1, 1, 0, 2>, "sp NCHW   softmax  write   noatom", chunk, 296, smem, bytes_w);
  for (int chunk : {1, 4, 64}) run(k_stream<0, 0, 0, 0, 2>, "sp NCHW   trivial  nowrite noatom", chunk, 296, smem, bytes_in);
  // trivial work allows 3 CTAs/SM (registers): more bytes in flight
  run(k_stream<0, 0, 0, 0, 1>, "sp NCHW   trivial  nowrite noatom (1 CTA/SM)", 8, 148, smem, bytes_in);
  run(k_stream_wide<0>, "wide(512thr) NCHW trivial", 4, 148, 16 * kC * 512, bytes_in, 512);
  run(k_stream_wide<1>, "wide(512thr) NCHW softmax", 4, 148, 16 * kC * 512, bytes_in, 512);
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
