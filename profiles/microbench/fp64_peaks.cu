// Round-1 microbenchmark: what FP64 rate can a B200 SM actually sustain?
//   (a) DMMA.8x8x4 issue rate (register-resident operands, NACC independent accumulators per warp)
//   (b) DFMA issue rate (plain FP64 FMA pipe)
//   (c) DMMA fed from shared memory with LDS.64 fragment loads (the real inner loop shape)
// Built with: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peaks fp64_peaks.cu
// Output: one JSON object on stdout.  Numbers feed DESIGN.md (FP64 roofline denominator next to cuBLAS DGEMM).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(1024) dmma_reg(double* out, const double* in, int iters) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) dfma_reg(double* out, const double* in, int iters) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(a, c[i], b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// warp tile 64x32 (8x4 DMMA tiles), k-step 4: per k-step 8 A-frag LDS.64 + 4 B-frag LDS.64, 32 DMMA.
// smem tile layout [k][m] with row pitch chosen conflict-free; this measures LDS+DMMA co-issue.
template <int WM, int WN>
__global__ void __launch_bounds__(256) dmma_smem(double* out, const double* in, int iters) {
  extern __shared__ double sm[];
  const int KT = 16;
  const int PA = 64 * 4 + 4;  // pitch (doubles) of A tile rows [k][m], m over 256 (4 warps x 64)
  const int PB = 32 * 2 + 4;
  double* sA = sm;             // [KT][PA]
  double* sB = sm + KT * PA;   // [KT][PB]
  for (int i = threadIdx.x; i < KT * PA + KT * PB; i += blockDim.x) sm[i] = in[i & 63];
  __syncthreads();
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int wm = (warp & 3) * 64, wn = (warp >> 2) * 32;
  int g = lane >> 2, t = lane & 3;
  double c[WM][WN][2];
#pragma unroll
  for (int i = 0; i < WM; i++)
#pragma unroll
    for (int j = 0; j < WN; j++) { c[i][j][0] = 0; c[i][j][1] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int kk = 0; kk < KT; kk += 4) {
      double a[WM], b[WN];
#pragma unroll
      for (int i = 0; i < WM; i++) a[i] = sA[(kk + t) * PA + wm + i * 8 + g];
#pragma unroll
      for (int j = 0; j < WN; j++) b[j] = sB[(kk + t) * PB + wn + j * 8 + g];
#pragma unroll
      for (int i = 0; i < WM; i++)
#pragma unroll
        for (int j = 0; j < WN; j++)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < WM; i++)
#pragma unroll
    for (int j = 0; j < WN; j++) s += c[i][j][0] + c[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f, int rep) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < rep; i++) f();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / rep;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double *in, *out; CK(cudaMalloc(&in, 1 << 20)); CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  CK(cudaMemset(in, 0, 1 << 20));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);
  int iters = 20000;
  // (a) DMMA from registers: threads/block x blocks/SM sweep
  int cfg[][2] = {{128, 1}, {256, 1}, {512, 1}, {1024, 1}, {256, 2}, {256, 4}};
  for (auto& c : cfg) {
    int th = c[0], bps = c[1];
    {
      float ms = time_ms([&] { dmma_reg<8><<<sms * bps, th>>>(out, in, iters); }, 3);
      double fl = 2.0 * 256 * 8 * (double)iters * (th / 32) * sms * bps;
      printf(", \"dmma_reg_acc8_t%d_b%d_tflops\": %.2f", th, bps, fl / ms / 1e9);
    }
    {
      float ms = time_ms([&] { dmma_reg<2><<<sms * bps, th>>>(out, in, iters); }, 3);
      double fl = 2.0 * 256 * 2 * (double)iters * (th / 32) * sms * bps;
      printf(", \"dmma_reg_acc2_t%d_b%d_tflops\": %.2f", th, bps, fl / ms / 1e9);
    }
    {
      float ms = time_ms([&] { dfma_reg<8><<<sms * bps, th>>>(out, in, iters); }, 3);
      double fl = 2.0 * 8 * (double)iters * th * sms * bps;
      printf(", \"dfma_reg_acc8_t%d_b%d_tflops\": %.2f", th, bps, fl / ms / 1e9);
    }
  }
  // (c) DMMA fed by LDS: 256 threads, warp tile 64x32
  {
    size_t smem = sizeof(double) * 16 * (260 + 68);
    int it2 = 4000;
    for (int bps = 1; bps <= 2; bps++) {
      float ms = time_ms([&] { dmma_smem<8, 4><<<sms * bps, 256, smem>>>(out, in, it2); }, 3);
      double fl = 2.0 * 256 * 32 * 4 * (double)it2 * 8 * sms * bps;
      printf(", \"dmma_smem_w64x32_b%d_tflops\": %.2f", bps, fl / ms / 1e9);
    }
    for (int bps = 1; bps <= 2; bps++) {
      float ms = time_ms([&] { dmma_smem<4, 4><<<sms * bps, 256, smem>>>(out, in, it2); }, 3);
      double fl = 2.0 * 256 * 16 * 4 * (double)it2 * 8 * sms * bps;
      printf(", \"dmma_smem_w32x32_b%d_tflops\": %.2f", bps, fl / ms / 1e9);
    }
  }
  printf("}\n");
  return 0;
}
