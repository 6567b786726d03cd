// fpv_benchmark -- the reference's benchmark on the GPU path: times
// Encoder::Init / CompressFrame x N / Finish over a raw file held in memory
// (benchmark.cc:153-180), prints the same figures (bytes, bpp, MP/s, fps) plus
// raw-pixel GB/s, then verifies the round trip through both decoders
// (benchmark.cc:192-285) and times the streaming decode as well.
// Command line as there: file xsize ysize big_endian shift [maxframes] [threads] [batch].
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <iostream>
#include <vector>

#include "../host/fusion_power_video.h"

static double Now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char* argv[]) {
  if (argc < 6) {
    std::cerr << "usage: " << argv[0] << " file xsize ysize big_endian shift [maxframes] [threads=8] [batch=32] [gpu_entropy=0] [gpus=1]\n";
    return 1;
  }
  const size_t xsize = strtoull(argv[2], nullptr, 10), ysize = strtoull(argv[3], nullptr, 10);
  const bool big_endian = atoi(argv[4]) != 0;
  const int shift = atoi(argv[5]);
  size_t maxframes = argc > 6 ? strtoull(argv[6], nullptr, 10) : 0;
  const size_t threads = argc > 7 ? strtoull(argv[7], nullptr, 10) : 8;
  fpvc::GpuOptions opt;
  if (argc > 8) opt.batch = (uint32_t)atoi(argv[8]);
  if (argc > 9) opt.gpu_entropy = atoi(argv[9]) != 0;
  if (argc > 10)
    for (int d = 0; d < atoi(argv[10]); d++) opt.devices.push_back(d);
  const size_t px = xsize * ysize;
  if (px == 0) return 1;

  FILE* f = fopen(argv[1], "rb");
  if (!f) { std::cerr << "cannot open " << argv[1] << "\n"; return 1; }
  std::vector<uint16_t> raw;
  std::vector<uint16_t> one(px);
  while ((maxframes == 0 || raw.size() / px < maxframes) && fread(one.data(), 2, px, f) == px)
    raw.insert(raw.end(), one.begin(), one.end());
  fclose(f);
  const size_t nframes = raw.size() / px;
  if (nframes == 0) { std::cerr << "no complete frame in file\n"; return 1; }

  std::vector<uint8_t> stream;
  auto append = [](const uint8_t* d, size_t n, void* p) {
    auto* v = static_cast<std::vector<uint8_t>*>(p);
    v->insert(v->end(), d, d + n);
  };
  const double t0 = Now();
  {
    fpvc::Encoder enc(threads, shift, big_endian, opt);
    enc.Init(raw.data(), xsize, ysize, append, &stream);  // frame 0 is the delta frame (benchmark.cc:146)
    if (!enc.ok()) { std::cerr << "encoder failed: " << fpvc::LastError() << "\n"; return 1; }
    for (size_t i = 0; i < nframes; i++) enc.CompressFrame(raw.data() + i * px, append, &stream);
    enc.Finish(append, &stream);
  }
  const double t_enc = Now() - t0;
  const double mp = (double)nframes * px / 1e6;
  printf("encode: %zu frames, %zu bytes, %.3f bpp, %.1f bytes/frame, %.2f ms, %.2f MP/s, %.2f fps, %.3f GB/s raw\n",
         nframes, stream.size(), stream.size() * 8.0 / (nframes * (double)px), stream.size() / (double)nframes,
         t_enc * 1e3, mp / t_enc, nframes / t_enc, nframes * px * 2.0 / t_enc / 1e9);

  // streaming decoder in 64 KiB blocks, compared after UnextractFrame (here: on the GPU)
  size_t decoded = 0, mismatches = 0;
  bool failed = false;
  std::vector<uint8_t> expect(px * 2);
  const double t1 = Now();
  {
    fpvc::StreamingDecoder dec(opt);
    dec.SetRawOutput(shift, big_endian);
    for (size_t pos = 0; pos < stream.size() && !failed; pos += 65536) {
      const size_t n = std::min<size_t>(65536, stream.size() - pos);
      dec.Decode(stream.data() + pos, n,
                 [&](bool ok, uint16_t* frame, size_t, size_t, void*) {
                   if (!ok) { failed = true; return; }
                   if (decoded < nframes && memcmp(frame, raw.data() + decoded * px, px * 2) != 0) mismatches++;
                   decoded++;
                 },
                 nullptr);
    }
  }
  const double t_dec = Now() - t1;
  // the comparison above is against the file bytes; they equal raw[] on a little-endian host
  printf("streaming decode: %zu frames, %.2f ms, %.2f MP/s, %s\n", decoded, t_dec * 1e3, mp / t_dec,
         (!failed && decoded == nframes && mismatches == 0) ? "ok" : "MISMATCH");

  fpvc::RandomAccessDecoder rad(opt);
  bool ra_ok = rad.Init(stream.data(), stream.size()) && rad.numframes() == nframes;
  std::vector<uint16_t> img(px);
  std::vector<uint8_t> preview((xsize / 4) * (ysize / 4) + 1), file_bytes(px * 2);
  for (size_t i = 0; ra_ok && i < nframes; i += (nframes > 16 ? nframes / 16 : 1)) {
    ra_ok = rad.DecodeFrame(i, img.data()) && rad.DecodePreview(i, preview.data());
    if (ra_ok) {
      fpvc::UnextractFrame(img.data(), xsize, ysize, shift, big_endian, file_bytes.data());
      ra_ok = memcmp(file_bytes.data(), raw.data() + i * px, px * 2) == 0;
    }
  }
  printf("random access decode: %s\n", ra_ok ? "ok" : "MISMATCH");
  return (!failed && decoded == nframes && mismatches == 0 && ra_ok) ? 0 : 1;
}
