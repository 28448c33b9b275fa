// Micro-benchmark of the single-CTA radix sort (diagnostic; built by hand: see the nvcc line in tools/probe/README).
#include "../../gpscalibration_b200/csrc/lg_voxel.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>
char g_cuda_err[512];
thread_local LgProf* g_lg_prof = nullptr;
void lg_set_error(const char* msg, const char* file, int line) { snprintf(g_cuda_err, sizeof(g_cuda_err), "%s (%s:%d)", msg, file, line); }
int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 10000, bits = argc > 2 ? atoi(argv[2]) : 32;
  cudaStream_t st;
  cudaStreamCreate(&st);
  RadixWs ws;
  lg_radix_ensure(ws, n, st);
  std::vector<unsigned long long> k(n);
  std::vector<unsigned int> v(n);
  srand(1);
  for (int i = 0; i < n; i++) { k[i] = ((unsigned long long)rand() << 16) ^ rand(); k[i] &= (bits >= 64 ? ~0ull : ((1ull << bits) - 1)); v[i] = i; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  long long launches = 0;
  float best = 1e9f;
  for (int rep = 0; rep < 20; rep++) {
    cudaMemcpyAsync(ws.keysA.p, k.data(), n * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(ws.valsA.p, v.data(), n * 4, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    int in_b = 0;
    cudaEventRecord(e0, st);
    lg_radix_sort(ws, n, bits, st, &launches, &in_b);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
    if (rep == 19) {
      std::vector<unsigned long long> out(n);
      cudaMemcpy(out.data(), in_b ? ws.keysB.p : ws.keysA.p, n * 8, cudaMemcpyDeviceToHost);
      std::vector<unsigned int> outv(n);
      cudaMemcpy(outv.data(), in_b ? ws.valsB.p : ws.valsA.p, n * 4, cudaMemcpyDeviceToHost);
      bool ok = true;
      for (int i = 1; i < n; i++) ok = ok && (out[i - 1] < out[i] || (out[i - 1] == out[i] && outv[i - 1] < outv[i]));
      for (int i = 0; i < n; i++) ok = ok && out[i] == k[outv[i]];
      printf("n=%d bits=%d best %.1f us sorted=%d err=%s\n", n, bits, best * 1e3f, (int)ok, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
