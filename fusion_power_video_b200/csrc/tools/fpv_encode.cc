// fpv_encode -- raw 16-bit frames on stdin -> fusion-power-video stream on stdout.
// Command line as the reference's encode (encode.cc:41-48; note the order it
// actually parses): xsize ysize big_endian shift [threads] [batch] [gpu_entropy] [gpus].
// The first frame doubles as the delta frame (encode.cc:87-90).
#include <stdio.h>
#include <stdlib.h>

#include <iostream>
#include <vector>

#include "../host/fusion_power_video.h"

int main(int argc, char* argv[]) {
  if (argc < 5) {
    std::cerr << "usage: " << argv[0] << " xsize ysize big_endian shift [threads=4] [batch=32] [gpu_entropy=0] [gpus=1] < raw > stream\n";
    return 1;
  }
  const size_t xsize = strtoull(argv[1], nullptr, 10), ysize = strtoull(argv[2], nullptr, 10);
  const bool big_endian = atoi(argv[3]) != 0;
  const int shift = atoi(argv[4]);
  const size_t threads = argc > 5 ? strtoull(argv[5], nullptr, 10) : 4;
  fpvc::GpuOptions opt;
  if (argc > 6) opt.batch = (uint32_t)atoi(argv[6]);
  if (argc > 7) opt.gpu_entropy = atoi(argv[7]) != 0;   // 1: entropy-code on the GPU (valid brotli, decodable by the reference)
  if (argc > 8)
    for (int d = 0; d < atoi(argv[8]); d++) opt.devices.push_back(d);   // one Encoder over several GPUs, same stream bytes
  if (xsize == 0 || xsize > 65536 || ysize == 0 || ysize > 65536 || shift < 0 || shift > 16) {
    std::cerr << "invalid arguments\n";
    return 1;
  }
  fpvc::Encoder encoder(threads, shift, big_endian, opt);
  auto write = [](const uint8_t* data, size_t size, void*) { fwrite(data, 1, size, stdout); };
  std::vector<uint16_t> frame(xsize * ysize);
  bool initialised = false;
  while (fread(frame.data(), 2, frame.size(), stdin) == frame.size()) {
    if (!initialised) {
      encoder.Init(frame.data(), xsize, ysize, write, nullptr);
      if (!encoder.ok()) {
        std::cerr << "encoder initialisation failed: " << fpvc::LastError() << "\n";
        return 1;
      }
      initialised = true;
    }
    encoder.CompressFrame(frame.data(), write, nullptr);  // copied before it returns
  }
  encoder.Finish(write, nullptr);
  return encoder.ok() || !initialised ? 0 : 1;
}
