// fpv_decode -- fusion-power-video stream on stdin -> raw 16-bit frames on stdout.
// Command line as the reference's decode (decode.cc:41-44): xsize ysize
// big_endian shift (xsize / ysize are taken from the stream, as there).
// UnextractFrame runs fused in the GPU kernel (StreamingDecoder::SetRawOutput).
#include <stdio.h>
#include <stdlib.h>

#include <iostream>
#include <vector>

#include "../host/fusion_power_video.h"

int main(int argc, char* argv[]) {
  if (argc < 5) {
    std::cerr << "usage: " << argv[0] << " xsize ysize big_endian shift [batch=32] < stream > raw\n";
    return 1;
  }
  const bool big_endian = atoi(argv[3]) != 0;
  const int shift = atoi(argv[4]);
  fpvc::GpuOptions opt;
  opt.batch = 32;
  if (argc > 5) opt.batch = (uint32_t)atoi(argv[5]);
  fpvc::StreamingDecoder decoder(opt);
  decoder.SetRawOutput(shift, big_endian);
  bool failed = false;
  // decode.cc reads 1 MiB at a time (decode.cc:67-77); the GPU decoder wants several batches of frames per call,
  // so that brotli decoding of one batch overlaps the inverse transform of the previous one
  std::vector<uint8_t> block((size_t)128 << 20);
  size_t n;
  while (!failed && (n = fread(block.data(), 1, block.size(), stdin)) > 0) {
    decoder.Decode(block.data(), n,
                   [&failed](bool ok, uint16_t* frame, size_t xs, size_t ys, void*) {
                     if (!ok) { failed = true; return; }
                     fwrite(frame, 2, xs * ys, stdout);
                   },
                   nullptr);
  }
  if (failed) std::cerr << "decoding failed: " << fpvc::LastError() << "\n";
  return failed ? 1 : 0;
}
