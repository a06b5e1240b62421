// Probe: which way of issuing a TMA tile load is legal on this GPU/driver (run on the B200 box).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include "../../pl-slam_b200/csrc/tma.cuh"
namespace pl { namespace tma {
// uint8 tensor [d2][d1][d0] with byte pitches p1 (rows) and p2 (frames); box = tile fetched per load.
// Requirements of the hardware: base and pitches multiples of 16 bytes, box0 a multiple of 16 bytes, every box side <= 256.
inline bool encode_u8_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t p1, uint64_t p2,
                         uint32_t box0, uint32_t box1) {
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                         const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
    fn = (Fn)p;
  }
  if (d2 == 1 && p2 < p1 * d1) p2 = p1 * d1;        // a single frame: the frame pitch is never used, but it has to be a legal one
  if (((uintptr_t)base & 15) || (p1 & 15) || (p2 & 15) || (box0 & 15) || box0 > 256 || box1 > 256 || p1 < d0 || p2 < p1 * d1) return false;
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {p1, p2};
  const cuuint32_t box[3] = {box0, box1, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// tile (c0, c1, c2) of a 3-D tensor map -> shared memory; completes `bar` with the tile's byte count
__device__ __forceinline__ void load_3d(void* smem_dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
} }
using namespace pl;
struct Maps { CUtensorMap m[4]; int bw, bh; };

template <int MODE>
__global__ void k(const __grid_constant__ Maps T, const __grid_constant__ CUtensorMap one, int level, unsigned* out, const CUtensorMap* gmap, const uint8_t* gsrc = nullptr) {
  __shared__ __align__(128) uint8_t win[64 * 64];
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x;
  if (tid == 0) { tma::mbar_init(&mbar, 1); tma::fence_mbar_init(); }
  __syncthreads();
  if (MODE == 0) { if (tid == 0) { tma::mbar_expect_tx(&mbar, T.bw * T.bh); tma::load_3d(win, &one, &mbar, 3, 5, 0); } }
  if (MODE == 1) { if (tid == 0) { tma::mbar_expect_tx(&mbar, T.bw * T.bh); tma::load_3d(win, &T.m[1], &mbar, 3, 5, 0); } }
  if (MODE == 2) { if (tid == 0) { tma::mbar_expect_tx(&mbar, T.bw * T.bh); tma::load_3d(win, &T.m[level], &mbar, 3, 5, 0); } }
  if (MODE == 4) { if (tid == 0) { tma::mbar_expect_tx(&mbar, T.bw * T.bh); tma::load_3d(win, gmap, &mbar, 3, 5, 0); } }
  if (MODE == 5) { if (tid == 0) { tma::mbar_expect_tx(&mbar, T.bw * T.bh); asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(tma::smem_u32(win)), "l"(reinterpret_cast<uint64_t>(gmap)), "r"(tma::smem_u32(&mbar)), "r"(3), "r"(5) : "memory"); } }
  if (MODE == 6) { if (tid == 0) { tma::mbar_expect_tx(&mbar, 256); asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tma::smem_u32(win)), "l"(gsrc), "r"(256), "r"(tma::smem_u32(&mbar)) : "memory"); } }
  if (MODE == 7) { if (tid == 0) { tma::mbar_expect_tx(&mbar, T.bw * T.bh); asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(tma::smem_u32(win)), "l"(reinterpret_cast<uint64_t>(gmap)), "r"(tma::smem_u32(&mbar)), "r"(3), "r"(5) : "memory"); } }
  if (MODE == 3) { if (threadIdx.x < 32 && __shfl_sync(0xffffffffu, 0, 0) == 0) { unsigned leader = 0; asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(leader)); if (leader) { tma::mbar_expect_tx(&mbar, T.bw * T.bh); tma::load_3d(win, &T.m[level], &mbar, 3, 5, 0); } } }
  tma::mbar_wait(&mbar, 0);
  if (tid == 0) { out[0] = win[0]; out[1] = win[T.bw + 1]; }
}
int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int W = 640, H = 480, B = 2;
  uint8_t* d; cudaMalloc(&d, (size_t)W * H * B);
  uint8_t* h = (uint8_t*)malloc((size_t)W * H * B);
  for (size_t i = 0; i < (size_t)W * H * B; i++) h[i] = (uint8_t)(i * 7 + (i >> 9));
  cudaMemcpy(d, h, (size_t)W * H * B, cudaMemcpyHostToDevice);
  Maps T; memset(&T, 0, sizeof(T)); T.bw = 48; T.bh = 37;
  bool ok = true;
  for (int l = 0; l < 4; l++) ok &= tma::encode_u8_3d(&T.m[l], d, W, H, B, W, (uint64_t)W * H, 48, 37);
  printf("encode ok=%d sizeof(CUtensorMap)=%zu alignof=%zu\n", (int)ok, sizeof(CUtensorMap), alignof(CUtensorMap));
  unsigned* out; cudaMalloc(&out, 8); cudaMemset(out, 0, 8);
  CUtensorMap* gmap; cudaMalloc(&gmap, 256);
  const int bw = argc > 2 ? atoi(argv[2]) : 48;
  if (mode == 4) { CUtensorMap m; ok = tma::encode_u8_3d(&m, d, W, H, B, W, (uint64_t)W * H, bw, 37); T.bw = bw; cudaMemcpy(gmap, &m, 128, cudaMemcpyHostToDevice); printf("global map, bw=%d ok=%d\n", bw, (int)ok); }
  if (mode == 5 || mode == 7) {
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* p = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap m; const cuuint64_t dims[2] = {640, 480}; const cuuint64_t st[1] = {640}; const cuuint32_t box[2] = {(cuuint32_t)bw, 37}; const cuuint32_t es[2] = {1, 1};
    CUresult r = ((Fn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    T.bw = bw; cudaMemcpy(gmap, &m, 128, cudaMemcpyHostToDevice); printf("2d global map bw=%d r=%d\n", bw, (int)r); }
  if (mode == 0) k<0><<<1, 128>>>(T, T.m[0], 1, out, gmap);
  if (mode == 1) k<1><<<1, 128>>>(T, T.m[0], 1, out, gmap);
  if (mode == 2) k<2><<<1, 128>>>(T, T.m[0], 1, out, gmap);
  if (mode == 3) k<3><<<1, 128>>>(T, T.m[0], 1, out, gmap);
  if (mode == 4) k<4><<<1, 128>>>(T, T.m[0], 1, out, gmap);
  if (mode == 5) k<5><<<1, 128>>>(T, T.m[0], 1, out, gmap);
  if (mode == 6) k<6><<<1, 128>>>(T, T.m[0], 1, out, gmap, d);
  if (mode == 8 || mode == 9) {   // 2-D map in global memory, launched WITH a cluster attribute (8) / with a uint32 element type (9)
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* p = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q);
    CUtensorMap m; CUresult r;
    if (mode == 8) { const cuuint64_t dims[2] = {640, 480}; const cuuint64_t st[1] = {640}; const cuuint32_t box[2] = {48, 37}; const cuuint32_t es[2] = {1, 1};
      r = ((Fn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    else { const cuuint64_t dims[2] = {160, 480}; const cuuint64_t st[1] = {640}; const cuuint32_t box[2] = {12, 37}; const cuuint32_t es[2] = {1, 1};
      r = ((Fn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    T.bw = 48; cudaMemcpy(gmap, &m, 128, cudaMemcpyHostToDevice); printf("mode %d map r=%d q=%d\n", mode, (int)r, (int)q);
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(1); cfg.blockDim = dim3(128);
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = (mode == 8) ? 1 : 0;
    const uint8_t* np = nullptr;
    cudaError_t le = cudaLaunchKernelEx(&cfg, k<5>, T, T.m[0], 1, out, (const CUtensorMap*)gmap, np);
    printf("launch: %s\n", cudaGetErrorString(le));
  }
  if (mode == 7) k<7><<<1, 128>>>(T, T.m[0], 1, out, gmap);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned r[2] = {0, 0}; cudaMemcpy(r, out, 8, cudaMemcpyDeviceToHost);
  printf("mode %d: %s  got %u %u expect %u %u\n", mode, cudaGetErrorString(e), r[0], r[1], (unsigned)h[5 * W + 3], (unsigned)h[6 * W + 4]);
  return 0;
}
