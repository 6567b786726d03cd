// tsan_driver.cc -- TEST INFRASTRUCTURE.  Runs fpvc::Encoder (one and several devices) many times on the
// oracle-backed stand-in under ThreadSanitizer and checks that every run writes the same bytes:
//     make -C tests/cpu_cabi tsan
// (round 2 found a race in Encoder::compress_batch this way: the loop re-read Batch::n after the batch had
// been recycled.)
#include <stdint.h>
#include <stdio.h>

#include <vector>

extern "C" size_t fpvh_encode_stream_multi(size_t xsize, size_t ysize, int shift, int big_endian, size_t threads,
                                           uint32_t batch, const int* devices, int ndevices, int gpu_entropy,
                                           const uint16_t* delta, const uint16_t* frames, size_t nframes, uint8_t* out,
                                           size_t cap, double* seconds);

int main() {
  const size_t W = 128, H = 64, n = 41, P = W * H;
  std::vector<uint16_t> fr(n * P);
  uint32_t s = 1;
  for (auto& v : fr) { s = s * 1664525u + 1013904223u; v = (uint16_t)(((s >> 12) & 0xfff) + 3000); }
  std::vector<uint8_t> out(n * P * 3), first;
  int devs[4] = {0, 1, 2, 3}, bad = 0;
  for (int it = 0; it < 40; it++) {
    const int nd = 1 + it % 4;
    const size_t sz = fpvh_encode_stream_multi(W, H, 0, 0, 4, 1 + it % 5, devs, nd, 0, fr.data(), fr.data(), n, out.data(),
                                               out.size(), nullptr);
    std::vector<uint8_t> cur(out.begin(), out.begin() + sz);
    if (it == 0) first = cur;
    else if (cur != first) { printf("iteration %d (%d devices) differs: %zu vs %zu bytes\n", it, nd, sz, first.size()); bad++; }
  }
  printf(bad ? "FAILED\n" : "all runs identical\n");
  return bad ? 1 : 0;
}
