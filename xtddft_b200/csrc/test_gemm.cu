// Standalone validation + timing of the DMMA/TMA GEMM against the naive checker kernel.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o test_gemm test_gemm.cu
#include "gemm.cuh"
#include <vector>
#include <cmath>
namespace xtd { thread_local char g_last_error[512]; unsigned long long g_launch_count = 0; }
using namespace xtd;

static void fill(std::vector<double>& v, unsigned seed) {
  unsigned long long s = seed * 2654435761ull + 12345;
  for (auto& x : v) { s = s * 6364136223846793005ull + 1442695040888963407ull; x = ((double)(s >> 11) / 9007199254740992.0) - 0.5; }
}
#define CHECK(x) do { int r_ = (x); if (r_ != 0) { printf("FAIL %s -> %d : %s\n", #x, r_, g_last_error); return 1; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s : %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

struct Case { int M, N, K, nouter, batches; bool akc, bkc; int row0, col0; int splits; bool acc; };

int main(int argc, char** argv) {
  GemmContext ctx;
  CHECK(gemm_context_init(ctx));
  CU(cudaMalloc(&ctx.split_ws, 1ull << 30)); ctx.split_ws_bytes = 1ull << 30;
  std::vector<Case> cases = {
    {128, 128, 16, 1, 1, true, true, 0, 0, 1, false},
    {128, 128, 64, 1, 1, true, true, 0, 0, 1, false},
    {100, 77, 45, 1, 1, true, true, 0, 0, 1, false},
    {277, 1777, 277, 1, 3, true, true, 0, 0, 1, false},
    {300, 200, 130, 3, 2, true, true, 5, 6, 1, true},
    {300, 200, 130, 3, 2, true, true, 5, 8, 1, true},
    {300, 200, 130, 3, 2, true, true, 0, 10, 1, false},
    {300, 200, 130, 3, 2, true, true, 5, 0, 1, false},
    {300, 200, 130, 1, 1, true, true, 0, 2, 1, false},
    {300, 200, 130, 3, 2, false, false, 5, 6, 1, false},
    {300, 200, 130, 2, 2, true, false, 3, 2, 1, false},
    {300, 200, 130, 2, 2, false, true, 3, 2, 1, true},
    {277, 555, 1000, 7, 1, true, true, 0, 0, 0, false},
    {50, 60, 5000, 1, 1, false, false, 0, 0, 0, true},
    {10, 48, 20000, 1, 1, true, true, 0, 0, 0, false},
    {1, 1, 1, 1, 1, true, true, 0, 0, 1, false},
  };
  int bad = 0;
  size_t c_lo = argc > 1 ? atoi(argv[1]) : 0, c_hi = argc > 2 ? atoi(argv[2]) : cases.size();
  bool do_time = argc <= 1 || (argc > 3 && atoi(argv[3]));
  for (size_t ci = c_lo; ci < c_hi && ci < cases.size(); ++ci) {
    Case c = cases[ci];
    int nq = c.nouter + 2 * c.batches;       // slices
    int arows = (c.akc ? c.M : c.K) + c.row0, acols = (c.akc ? c.K : c.M) + c.col0;
    int brows = (c.bkc ? c.N : c.K) + c.row0, bcols = (c.bkc ? c.K : c.N) + c.col0;
    long lda = pad_ld(acols + 3), ldb = pad_ld(bcols + 5), ldc = pad_ld(c.N);
    long sqa = lda * (arows + 2), sqb = ldb * (brows + 1);
    std::vector<double> hA(sqa * nq), hB(sqb * nq), hC((long)c.batches * c.M * ldc);
    fill(hA, 1 + ci); fill(hB, 100 + ci); fill(hC, 200 + ci);
    double *dA, *dB, *dC1, *dC2;
    CU(cudaMalloc(&dA, hA.size() * 8)); CU(cudaMalloc(&dB, hB.size() * 8));
    CU(cudaMalloc(&dC1, hC.size() * 8)); CU(cudaMalloc(&dC2, hC.size() * 8));
    CU(cudaMemcpy(dA, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dB, hB.data(), hB.size() * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dC1, hC.data(), hC.size() * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dC2, hC.data(), hC.size() * 8, cudaMemcpyHostToDevice));
    GemmDesc d;
    d.A.base = dA; d.A.ld = lda; d.A.sq = sqa; d.A.row0 = c.row0; d.A.col0 = c.col0;
    d.A.rows = c.akc ? c.M : c.K; d.A.cols = c.akc ? c.K : c.M; d.A.q0 = 0; d.A.nq = nq;
    d.B.base = dB; d.B.ld = ldb; d.B.sq = sqb; d.B.row0 = c.row0; d.B.col0 = c.col0;
    d.B.rows = c.bkc ? c.N : c.K; d.B.cols = c.bkc ? c.K : c.N; d.B.q0 = 1; d.B.nq = nq - 1;
    d.a_kc = c.akc; d.b_kc = c.bkc; d.M = c.M; d.N = c.N; d.K = c.K; d.nouter = c.nouter; d.batches = c.batches;
    d.z_div = 1; d.a_hi = 1; d.b_hi = 1;
    d.C = dC1; d.ldc = ldc; d.c_batch_stride = (long)c.M * ldc; d.alpha = 0.75; d.accumulate = c.acc; d.splits = c.splits;
    ctx.naive = false;
    CHECK(gemm(ctx, d, 0));
    CU(cudaDeviceSynchronize());
    ctx.naive = true; d.C = dC2;
    CHECK(gemm(ctx, d, 0));
    CU(cudaDeviceSynchronize());
    std::vector<double> r1(hC.size()), r2(hC.size());
    CU(cudaMemcpy(r1.data(), dC1, hC.size() * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(r2.data(), dC2, hC.size() * 8, cudaMemcpyDeviceToHost));
    double maxd = 0, maxv = 0;
    for (size_t i = 0; i < r1.size(); ++i) { maxd = fmax(maxd, fabs(r1[i] - r2[i])); maxv = fmax(maxv, fabs(r2[i])); }
    bool ok = maxd <= 1e-12 * fmax(1.0, maxv) * sqrt((double)c.K * c.nouter);
    printf("case %zu M%d N%d K%d outer%d batch%d %s%s off(%d,%d) splits%d acc%d : maxdiff %.3e (max %.3e) %s\n", ci, c.M, c.N, c.K,
           c.nouter, c.batches, c.akc ? "K" : "M", c.bkc ? "K" : "M", c.row0, c.col0, c.splits, (int)c.acc, maxd, maxv, ok ? "OK" : "BAD");
    if (!ok) bad++;
    cudaFree(dA); cudaFree(dB); cudaFree(dC1); cudaFree(dC2);
  }
  // ---- timing: shapes of the sigma path --------------------------------------------------------------
  struct T { int M, N, K, nouter; bool akc, bkc; const char* name; };
  std::vector<T> ts = {
    {8192, 8192, 8192, 1, true, true, "square8192 TN"},
    {4096, 4096, 4096, 1, true, true, "square4096 TN"},
    {2770, 1777, 1777, 64, true, true, "K2 x=10 (U*Lvv, 64 aux)"},
    {277, 1777, 1777, 256, true, true, "K2 x=1 (256 aux, split)"},
    {65536, 277, 2052, 1, true, true, "G1 ao*mo1 x=1"},
    {65536, 2770, 2052, 1, true, true, "G1 ao*mo1 x=10"},
    {2770, 2052, 65536, 1, false, false, "G2 A^T*ao x=10"},
    {277, 2052, 65536, 1, false, false, "G2 A^T*ao x=1 (split)"},
  };
  ctx.naive = false;
  if (do_time) for (auto& t : ts) {
    long arows = t.akc ? t.M : t.K, acols = t.akc ? t.K : t.M, brows = t.bkc ? t.N : t.K, bcols = t.bkc ? t.K : t.N;
    long lda = pad_ld(acols), ldb = pad_ld(bcols), ldc = pad_ld(t.N);
    long sqa = lda * arows, sqb = ldb * brows;
    double *dA, *dB, *dC;
    CU(cudaMalloc(&dA, sqa * t.nouter * 8)); CU(cudaMalloc(&dB, sqb * t.nouter * 8)); CU(cudaMalloc(&dC, (long)t.M * ldc * 8));
    CU(cudaMemset(dA, 0, sqa * t.nouter * 8)); CU(cudaMemset(dB, 0, sqb * t.nouter * 8));
    GemmDesc d;
    d.A.base = dA; d.A.ld = lda; d.A.sq = sqa; d.A.rows = (int)arows; d.A.cols = (int)acols; d.A.nq = t.nouter;
    d.B.base = dB; d.B.ld = ldb; d.B.sq = sqb; d.B.rows = (int)brows; d.B.cols = (int)bcols; d.B.nq = t.nouter;
    d.a_kc = t.akc; d.b_kc = t.bkc; d.M = t.M; d.N = t.N; d.K = t.K; d.nouter = t.nouter;
    d.C = dC; d.ldc = ldc; d.c_batch_stride = 0; d.splits = 0;
    CHECK(gemm(ctx, d, 0)); CHECK(gemm(ctx, d, 0));
    CU(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rep = 5;
    cudaEventRecord(e0);
    for (int r = 0; r < rep; ++r) CHECK(gemm(ctx, d, 0));
    cudaEventRecord(e1); CU(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= rep;
    double fl = 2.0 * t.M * t.N * (double)t.K * t.nouter;
    printf("time %-28s M%d N%d K%d outer%d : %.3f ms  %.2f TFLOP/s\n", t.name, t.M, t.N, t.K, t.nouter, ms, fl / ms / 1e9);
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
  }
  printf("%s\n", bad ? "SOME BAD" : "ALL OK");
  return bad;
}
