// Development microbenchmark: achievable HBM bandwidth of streaming kernels with R read
// streams and W write streams (16-byte vectors, grid-stride), to put the stage kernels'
// read/write mixes (stage 0: 2R:4W, stages 1-4: 4R:4W, tensor: 4R+src:4W) into perspective.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mb scripts/microbench_rw.cu && /tmp/mb
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

template <int R, int W, bool CS>
__global__ void __launch_bounds__(256) mix(const float4* const* __restrict__ in, float4* const* __restrict__ out,
                                           size_t n) {
  const float4* ip[R > 0 ? R : 1];
  float4* op[W > 0 ? W : 1];
#pragma unroll
  for (int r = 0; r < R; ++r) ip[r] = in[r];
#pragma unroll
  for (int w = 0; w < W; ++w) op[w] = out[w];
  float4 acc = make_float4(0, 0, 0, 0);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(1.f, 2.f, 3.f, 4.f);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float4 v = CS ? __ldcs(ip[r] + i) : ip[r][i];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
      float4 v = make_float4(s.x + w, s.y, s.z, s.w);
      if (CS) __stcs(op[w] + i, v); else op[w][i] = v;
    }
    if (W == 0) { acc.x += s.x; acc.y += s.y; acc.z += s.z; acc.w += s.w; }
  }
  if (W == 0 && acc.x == 123.456f) out[0][0] = acc;
}

template <int R, int W, bool CS>
void run(float4** bufs, const float4** d_in, float4** d_out, size_t n, int blocks) {
  const float4* hin[8]; float4* hout[8];
  for (int r = 0; r < R; ++r) hin[r] = bufs[r];
  for (int w = 0; w < (W > 0 ? W : 1); ++w) hout[w] = bufs[8 + w];
  cudaMemcpy(d_in, hin, sizeof(hin), cudaMemcpyHostToDevice);
  cudaMemcpy(d_out, hout, sizeof(hout), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) mix<R, W, CS><<<blocks, 256>>>(d_in, d_out, n);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    mix<R, W, CS><<<blocks, 256>>>(d_in, d_out, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double gb = (double)(R + W) * n * 16 / 1e9;
  printf("R=%d W=%d %s blocks=%d: %.3f ms  %.0f GB/s\n", R, W, CS ? "cs" : "plain", blocks, best,
         gb / best * 1e3);
}

int main() {
  const size_t n = (size_t)64 << 20;   // 64 Mi float4 = 1 GiB per stream
  float4* bufs[16];
  for (int i = 0; i < 16; ++i) { cudaMalloc(&bufs[i], n * 16); cudaMemset(bufs[i], 0, n * 16); }
  const float4** d_in; float4** d_out;
  cudaMalloc(&d_in, 64); cudaMalloc(&d_out, 64);
  for (int blocks : {148 * 8, 148 * 16}) {
    run<1, 1, false>(bufs, d_in, d_out, n, blocks);
    run<1, 1, true>(bufs, d_in, d_out, n, blocks);
    run<1, 0, true>(bufs, d_in, d_out, n, blocks);
    run<0, 1, true>(bufs, d_in, d_out, n, blocks);
    run<0, 1, false>(bufs, d_in, d_out, n, blocks);
    run<2, 4, true>(bufs, d_in, d_out, n, blocks);
    run<2, 4, false>(bufs, d_in, d_out, n, blocks);
    run<4, 4, true>(bufs, d_in, d_out, n, blocks);
    run<4, 4, false>(bufs, d_in, d_out, n, blocks);
    run<2, 1, true>(bufs, d_in, d_out, n, blocks);
    run<6, 4, true>(bufs, d_in, d_out, n, blocks);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
