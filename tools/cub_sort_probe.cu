// tools/cub_sort_probe.cu — NOT part of the product: times the CUDA toolkit's own radix sort (CUB
// DeviceRadixSort, a library kernel) on the C5 shape, as an outside reference point for the K5 numbers in
// profiles/ (how far is bci_sort from the vendor's implementation on the same box?).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/cub_sort_probe.cu -o build/tools/cub_sort_probe
#include <cub/cub.cuh>
#include <cstdio>
#include <vector>
#include <random>

int main() {
  const size_t n = 10000000, B = 32;
  float *in, *out;
  cudaMalloc(&in, n * B * 4);
  cudaMalloc(&out, n * B * 4);
  std::vector<float> h(n);
  std::mt19937 g(1);
  std::uniform_real_distribution<float> u(0.f, 1.f);
  for (auto& x : h) x = u(g);
  for (size_t b = 0; b < B; ++b) cudaMemcpy(in + b * n, h.data(), n * 4, cudaMemcpyHostToDevice);
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, in, out, (int)n);
  size_t big_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, big_bytes, in, out, (long long)(n * B));
  cudaMalloc(&tmp, big_bytes > tmp_bytes ? big_bytes : tmp_bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    for (size_t b = 0; b < B; ++b) cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, in + b * n, out + b * n, (int)n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cub 32 x 1e7 float keys, one call each: %.3f ms  %.2f Gkeys/s\n", ms, n * B / ms / 1e6);
  }
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    cub::DeviceRadixSort::SortKeys(tmp, big_bytes, in, out, (long long)(n * B));
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cub 3.2e8 float keys, one call: %.3f ms  %.2f Gkeys/s\n", ms, n * B / ms / 1e6);
  }
  return 0;
}
