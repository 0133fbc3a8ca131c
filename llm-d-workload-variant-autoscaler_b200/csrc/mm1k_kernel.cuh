// mm1k_kernel.cuh — closed-form M/M/1/K statistics, MM1KModel.Solve
// (pkg/analyzer/mm1kmodel.go:30-92, queuemodel.go:27-37), one thread per
// (lambda, mu, K) triple.  Outputs are float32 and are compared with the
// reference at 1e-6 relative (Go's math.Pow and CUDA's pow are not bit-identical,
// SURVEY.md §7 H2): rho^i is advanced by multiplication (error O(K) ulp of a
// float64, far inside the tolerance) with pow() only for p0.
#pragma once
#include "wva_core.cuh"

namespace wva {

__global__ void __launch_bounds__(128) mm1k_kernel(long long n, const float* lambda, const float* mu, const int* Kv,
                                                   unsigned char* valid, float* avg_resp, float* avg_wait,
                                                   float* avg_serv, float* avg_num, float* avg_queue,
                                                   float* throughput, float* rho_out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float lam = lambda[i], m = mu[i];
  int K = Kv[i];
  float rho = (lam == m) ? 1.0f : f_div(lam, m);                    // ComputeRho mm1kmodel.go:36-42
  bool ok = !((rho < 0.0f) || (rho >= (float)K) || (lam < 0.0f) || (m <= 0.0f));  // queuemodel.go:31
  float resp = 0, wait = 0, serv = 0, num = 0, queue = 0, tput = 0;
  if (ok) {
    double r = (double)rho;
    double p0 = (rho == 1.0f) ? 1.0 / (double)(K + 1) : (1.0 - r) / (1.0 - pow(r, (double)(K + 1)));
    double L = 0.0, pw = 1.0, pK = p0;
    for (int k = 1; k <= K; k++) {
      pw *= r;
      pK = p0 * pw;
      L += (double)k * pK;
    }
    num = (float)L;
    tput = f_mul(lam, f_sub(1.0f, (float)pK));
    resp = f_div(num, tput);
    serv = f_div(1.0f, m);
    wait = f_sub(resp, serv);
    if (wait < 0.0f) wait = 0.0f;
    queue = f_mul(tput, wait);
  }
  valid[i] = ok ? 1 : 0;
  avg_resp[i] = resp; avg_wait[i] = wait; avg_serv[i] = serv; avg_num[i] = num; avg_queue[i] = queue;
  throughput[i] = tput; rho_out[i] = rho;
}

}  // namespace wva
