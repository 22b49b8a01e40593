// Standalone probe for the TMA tile load used by k_orient_describe (debug aid).
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK, int FENCE>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int x, int y, int z, unsigned* out) {
  __shared__ __align__(128) uint8_t tile[1024];
  __shared__ __align__(8) unsigned long long mbar;
  const int lane = threadIdx.x;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    if (FENCE == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (FENCE == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  if (lane == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1024) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(smem_u32(tile)), "l"(&tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&mbar)) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(tile)), "l"(&tmap), "r"(x), "r"(y), "r"(smem_u32(&mbar)) : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
  __syncwarp();
  unsigned s = 0;
  for (int e = lane; e < 1024; e += 32) s += tile[e];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) *out = s;
}
using barrier_t = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
__global__ void ref_kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, unsigned* out) {
  __shared__ alignas(128) unsigned char smem_buffer[32][32];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier_t bar;
  if (threadIdx.x == 0) {
    init(&bar, blockDim.x);
    cde::fence_proxy_async_shared_cta();
  }
  __syncthreads();
  barrier_t::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  unsigned s = 0;
  for (int e = threadIdx.x; e < 1024; e += 32) s += ((unsigned char*)smem_buffer)[e];
  atomicAdd(out, s);
}
__global__ void bulk1d(const uint8_t* src, unsigned* out) {
  __shared__ __align__(128) uint8_t tile[1024];
  __shared__ __align__(8) unsigned long long mbar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1024) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(tile)), "l"(src), "r"(1024), "r"(smem_u32(&mbar)) : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
  __syncwarp();
  unsigned s = 0;
  for (int e = threadIdx.x; e < 1024; e += 32) s += tile[e];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) *out = s;
}
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int W = 752, H = 480, M = 4;
  std::vector<uint8_t> h((size_t)W * H * M);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 9));
  uint8_t* d;
  cudaMalloc(&d, h.size());
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  alignas(64) CUtensorMap tm;
  CUresult r;
  const int rank = (variant & 1) ? 2 : 3;
  if (rank == 3) {
    cuuint64_t dims[3] = {W, H, M}, str[2] = {W, (cuuint64_t)W * H};
    cuuint32_t box[3] = {32, 32, 1}, es[3] = {1, 1, 1};
    r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[2] = {W, (cuuint64_t)H * M}, str[1] = {W};
    cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
    r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  printf("variant %d rank %d encode=%d\n", variant, rank, (int)r);
  int drv = 0, rt = 0;
  cudaDriverGetVersion(&drv);
  cudaRuntimeGetVersion(&rt);
  printf("driver %d runtime %d; tensormap words:", drv, rt);
  for (int i = 0; i < 16; ++i) printf(" %016llx", ((unsigned long long*)&tm)[i]);
  printf("\n");
  unsigned* out;
  cudaMalloc(&out, 4);
  const int x = 100, y = 50, z = 1;
  const int fence = (variant >> 1) & 1;
  cudaMemset(out, 0, 4);
  if (variant == 5) {
    ref_kernel<<<1, 32>>>(tm, x, y + z * H, out);
  } else if (variant == 6) {
    bulk1d<<<1, 32>>>(d + 4096, out);
  } else
  { if (rank == 3 && fence == 0) probe<3, 0><<<1, 32>>>(tm, x, y, z, out);
  if (rank == 3 && fence == 1) probe<3, 1><<<1, 32>>>(tm, x, y, z, out);
  if (rank == 2 && fence == 0) probe<2, 0><<<1, 32>>>(tm, x, y + z * H, 0, out);
  if (rank == 2 && fence == 1) probe<2, 1><<<1, 32>>>(tm, x, y + z * H, 0, out); }
  cudaError_t e = cudaDeviceSynchronize();
  unsigned got = 0;
  cudaMemcpy(&got, out, 4, cudaMemcpyDeviceToHost);
  unsigned want = 0;
  for (int yy = 0; yy < 32; ++yy)
    for (int xx = 0; xx < 32; ++xx) want += h[((size_t)z * H + y + yy) * W + x + xx];
  printf("variant %d: %s got %u want %u\n", variant, cudaGetErrorString(e), got, want);
  return 0;
}
