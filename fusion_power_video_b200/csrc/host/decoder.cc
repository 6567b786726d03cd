// decoder.cc -- fpvc::StreamingDecoder and fpvc::RandomAccessDecoder on the GPU
// inverse transform.
//
// Both follow the reference's parsing and error behaviour
// (fusion_power_video.cc:866-956 and :961-1070) but split DecompressImage
// (.cc:296-347) in two: the brotli streams of all frames that are available are
// decoded on host threads straight into pinned plane buffers, then ONE
// fpv_decode call runs .cc:326-344 for the whole batch on the GPU.
#include <string.h>

#include <algorithm>
#include <future>
#include <thread>

#include "fusion_power_video.h"
#include "host_internal.h"

namespace fpvc {

using namespace internal;

namespace {

// Everything a decoder needs on the GPU side for one geometry.  Two sets of plane buffers: while the GPU
// inverts batch k and the caller's callbacks consume it, the pool already brotli-decodes batch k + 1.
struct GpuDecoder {
  fpv_ctx* ctx = nullptr;
  size_t W = 0, H = 0, P = 0;
  uint32_t B = 1;
  struct Set {
    Pinned high, low, flags;
    std::vector<char> good;
    size_t n = 0, left = 0;
    std::mutex m;
    std::condition_variable cv;
    // batches whose plane streams all carry the GPU coder's directories skip libbrotlidec: the coded bytes go
    // to the GPU as they are (fpv_decode_coded)
    bool coded = false;
    Pinned blob;
    size_t blob_bytes = 0;
    std::vector<fpv_coded_chunk> chunks;
    std::vector<size_t> core_at, core_size;    // where each frame's core sits in the blob (for the fallback)
  } sets[3];
  bool coded_ok = true;        // false once the C ABI said it cannot decode coded chunks (CPU stand-in)
  Pinned out, out1, out2;       // decoded frames of set 0 / 1 / 2
  std::unique_ptr<Pool> pool;

  Pinned& outbuf(int which) { return which == 0 ? out : which == 1 ? out1 : out2; }

  ~GpuDecoder() { close(); }

  // Back to the unopened state (a decoder that is re-initialised with another geometry starts over).
  void close() {
    pool.reset();   // no parse task may outlive the buffers
    if (ctx) fpv_destroy(ctx);
    ctx = nullptr;
    for (Set& st : sets) {
      st.high.reset();
      st.low.reset();
      st.flags.reset();
      st.blob.reset();
    }
    out.reset();
    out1.reset();
    out2.reset();
  }

  bool alloc_set(Set& st) {
    if (st.high.bytes()) return true;
    if (!st.high.alloc((size_t)B * P) || !st.low.alloc((size_t)B * P) || !st.flags.alloc(B))
      return FPV_FAIL("pinned allocation failed");
    return true;
  }

  bool open(const GpuOptions& opt, size_t xsize, size_t ysize, int shift, bool big_endian, uint32_t batch) {
    W = xsize;
    H = ysize;
    P = xsize * ysize;
    B = batch ? batch : 1;
    if (fpv_create(&ctx, opt.device, (uint32_t)xsize, (uint32_t)ysize, shift, big_endian ? 1 : 0, B) != FPV_OK)
      return FPV_FAIL(std::string("fpv_create: ") + fpv_last_error(nullptr));
    if (!alloc_set(sets[0]) || !out.alloc((size_t)B * P * 2)) return false;
    size_t t = std::thread::hardware_concurrency();
    t = std::min<size_t>(t ? t : 1, 32);
    if (B > 1 && t > 1) pool.reset(new Pool(std::min<size_t>(t, B)));
    return true;
  }

  // Starts the brotli decoding of n <= B core chunks into buffer set `which` (returns at once when there is
  // a pool).  cores / sizes must stay valid until finish().
  // Directory-carrying streams only: lays the cores of the batch out in the pinned blob and lists their chunks.
  bool scan_coded(Set& st, const uint8_t* const* cores, const size_t* sizes, size_t n, bool allow_delta) {
    if (!coded_ok || n == 0 || getenv("FPV_HOST_BROTLI_DECODE")) return false;
    st.chunks.clear();
    std::vector<uint64_t> offs;
    size_t total = 0;
    std::vector<size_t> base(n);
    for (size_t i = 0; i < n; i++) {
      if (sizes[i] < 2) return false;
      const uint8_t f = cores[i][0];
      if ((f & FPV_FLAG_USE_DELTA) && !allow_delta) return false;     // the brotli path reports it per frame
      size_t pos = 1;
      base[i] = total;
      for (int plane = (f & FPV_FLAG_NO_LOW_BYTES) ? 0 : 1; plane >= 0; plane--) {   // low comes first (.cc:658-662)
        offs.clear();
        size_t len = 0;
        if (!ScanCodedPlane(cores[i] + pos, sizes[i] - pos, P, &offs, &len)) return false;
        for (size_t k = 0; k < offs.size(); k++)
          st.chunks.push_back(fpv_coded_chunk{total + pos + offs[k], (uint32_t)i, (uint32_t)plane, (uint32_t)k, 0});
        pos += len;
      }
      if (pos != sizes[i]) return false;
      total += (sizes[i] + 15) & ~(size_t)15;
    }
    if (st.blob.bytes() < total && !st.blob.alloc(std::max(total, (size_t)B * (2 * P + 4096)))) return false;
    st.blob_bytes = total;
    st.core_at = base;
    st.core_size.assign(sizes, sizes + n);
    st.good.assign(n, 1);
    st.n = n;
    st.left = n;
    for (size_t i = 0; i < n; i++) {
      const uint8_t* core = cores[i];
      const size_t size = sizes[i], at = base[i];
      auto task = [this, &st, core, size, at, i] {
        memcpy(st.blob.as<uint8_t>() + at, core, size);
        st.flags.as<uint8_t>()[i] = core[0];
        std::lock_guard<std::mutex> l(st.m);
        if (--st.left == 0) st.cv.notify_all();
      };
      if (pool) pool->run(task);
      else task();
    }
    return true;
  }

  bool start(int which, const uint8_t* const* cores, const size_t* sizes, size_t n, bool allow_delta) {
    Set& st = sets[which];
    if (!alloc_set(st)) return false;
    if (which != 0 && !outbuf(which).bytes() && !outbuf(which).alloc((size_t)B * P * 2)) return FPV_FAIL("pinned allocation failed");
    st.coded = scan_coded(st, cores, sizes, n, allow_delta);
    if (st.coded) return true;
    st.good.assign(n, 0);
    st.n = n;
    st.left = n;
    for (size_t i = 0; i < n; i++) {
      const uint8_t* core = cores[i];
      const size_t size = sizes[i];
      auto task = [this, &st, core, size, i, allow_delta] {
        uint8_t f = 0;
        bool ok = ParseCore(core, size, P, &f, st.high.as<uint8_t>() + i * P, st.low.as<uint8_t>() + i * P);
        if (ok && (f & FPV_FLAG_USE_DELTA) && !allow_delta) ok = FPV_FAIL("delta frame not given");
        st.flags.as<uint8_t>()[i] = f;
        st.good[i] = ok ? 1 : 0;
        std::lock_guard<std::mutex> l(st.m);
        if (--st.left == 0) st.cv.notify_all();
      };
      if (pool) pool->run(task);
      else task();
    }
    return true;
  }

  // Waits for start(which), then inverts the leading frames that parsed correctly into `out` (uint16 images,
  // or raw file bytes with FPV_DEC_UNEXTRACT).  Returns their number (n on success).
  size_t finish(int which, uint32_t options) {
    Set& st = sets[which];
    {
      std::unique_lock<std::mutex> l(st.m);
      st.cv.wait(l, [&] { return st.left == 0; });
    }
    size_t k = 0;
    while (k < st.n && st.good[k]) k++;
    if (k == 0) return 0;
    if (st.coded) {
      const int rc = fpv_decode_coded(ctx, st.blob.as<uint8_t>(), st.blob_bytes, st.chunks.data(), (uint32_t)st.chunks.size(),
                                      st.flags.as<uint8_t>(), (uint32_t)k, options, outbuf(which).as<uint8_t>());
      if (rc == FPV_OK) return k;
      // A chunk the GPU decoder refused (or a C ABI without one): the same bytes through libbrotlidec, frame by
      // frame, so that malformed streams fail exactly like libbrotli-coded ones
      if (rc == FPV_ERR_UNSUPPORTED) coded_ok = false;
      st.coded = false;
      for (size_t i = 0; i < st.n; i++) {
        uint8_t f = 0;
        st.good[i] = ParseCore(st.blob.as<uint8_t>() + st.core_at[i], st.core_size[i], P, &f, st.high.as<uint8_t>() + i * P,
                               st.low.as<uint8_t>() + i * P) ? 1 : 0;
        st.flags.as<uint8_t>()[i] = f;
      }
      k = 0;
      while (k < st.n && st.good[k]) k++;
      if (k == 0) return 0;
    }
    if (fpv_decode(ctx, st.high.as<uint8_t>(), st.low.as<uint8_t>(), st.flags.as<uint8_t>(), (uint32_t)k, options,
                   outbuf(which).as<uint8_t>()) != FPV_OK) {
      FPV_FAIL(std::string("fpv_decode: ") + fpv_last_error(ctx));
      return 0;
    }
    return k;
  }

  size_t decode(const uint8_t* const* cores, const size_t* sizes, size_t n, uint32_t options, bool allow_delta) {
    if (!start(0, cores, sizes, n, allow_delta)) return 0;
    return finish(0, options);
  }
};

bool CheckDims(size_t xsize, size_t ysize) {
  if (xsize == 0 || ysize == 0) return FPV_FAIL("invalid image dimensions");
  if (xsize > 65536 || ysize > 65536 || xsize * ysize > kMaxPixels) return FPV_FAIL("image too large");
  return true;
}

}  // namespace

// =====================================================================================
// StreamingDecoder
// =====================================================================================

struct StreamingDecoder::Impl {
  GpuOptions opt;
  GpuDecoder gpu;
  std::vector<uint8_t> buffer;
  bool have_delta = false;
  bool raw_output = false;
  int shift = 0;
  bool big_endian = false;
  size_t id = 0;
};

StreamingDecoder::StreamingDecoder() : StreamingDecoder(GpuOptions()) {}
StreamingDecoder::StreamingDecoder(const GpuOptions& options) : impl_(new Impl) { impl_->opt = options; }
StreamingDecoder::~StreamingDecoder() = default;

void StreamingDecoder::SetRawOutput(int shift, bool big_endian) {
  if (impl_->gpu.ctx && (shift != impl_->shift || big_endian != impl_->big_endian || !impl_->raw_output)) {
    // shift and endianness are properties of the GPU context, which exists (and holds the delta frame) by now
    FPV_FAIL("StreamingDecoder::SetRawOutput must be called before the first Decode(); ignored");
    return;
  }
  impl_->raw_output = true;
  impl_->shift = shift;
  impl_->big_endian = big_endian;
}

void StreamingDecoder::Decode(
    const uint8_t* bytes, size_t size,
    std::function<void(bool ok, uint16_t* frame, size_t xsize, size_t ysize, void* payload)> callback,
    void* payload) {
  Impl& s = *impl_;
  // As the reference: only copy when bytes from an earlier call are pending.
  if (!s.buffer.empty()) s.buffer.insert(s.buffer.end(), bytes, bytes + size);
  const uint8_t* in = s.buffer.empty() ? bytes : s.buffer.data();
  const size_t insize = s.buffer.empty() ? size : s.buffer.size();
  auto fail = [&](const char* what) { callback(FPV_FAIL(what), nullptr, 0, 0, payload); };

  size_t pos = 0;
  if (!s.have_delta && insize > 13) {
    const size_t xsize = LoadU32(in), ysize = LoadU32(in + 4);
    if (!CheckDims(xsize, ysize)) return fail("invalid stream header");
    const size_t chunk = LoadU32(in + 8);
    if (chunk < 5) return fail("too small for delta frame");
    if (in[12] != kChunkDelta) return fail("not a delta frame");
    if (8 + chunk <= insize) {
      if (!s.gpu.ctx && !s.gpu.open(s.opt, xsize, ysize, s.shift, s.big_endian, s.opt.batch))
        return fail("no GPU decoder");
      const uint8_t* core = in + 13;
      const size_t core_size = chunk - 5;
      if (s.gpu.decode(&core, &core_size, 1, FPV_DEC_DEFAULT, false) != 1)
        return fail("decompressing delta frame failed");
      if (fpv_set_delta_image(s.gpu.ctx, s.gpu.out.as<uint16_t>()) != FPV_OK) return fail("fpv_set_delta_image");
      s.have_delta = true;
      pos = 8 + chunk;
    }
  }

  const uint32_t options = s.raw_output ? FPV_DEC_UNEXTRACT : FPV_DEC_DEFAULT;
  std::vector<const uint8_t*> cores[3];
  std::vector<size_t> sizes[3];
  bool stream_bad = false;
  const char* bad_what = nullptr;
  size_t scan = pos;
  // the frames that are complete in the buffer from `scan` on, up to one GPU batch; their parsing starts at once
  auto gather = [&](int set) {
    cores[set].clear();
    sizes[set].clear();
    while (cores[set].size() < s.gpu.B && !stream_bad) {
      if (scan + 9 > insize) break;
      const size_t frame_size = LoadU32(in + scan);
      const uint8_t flag = in[scan + 4];
      if (flag == kChunkIndex) break;  // frame index: end of frames
      if (flag != kChunkFrame) { stream_bad = true; bad_what = "not a standard frame"; break; }
      if (scan + frame_size > insize) break;
      const size_t preview_size = LoadU32(in + scan + 5);
      if (preview_size > frame_size || frame_size < preview_size + 9) {
        stream_bad = true;
        bad_what = "preview size too large";
        break;
      }
      cores[set].push_back(in + scan + 9 + preview_size);
      sizes[set].push_back(frame_size - preview_size - 9);
      scan += frame_size;
    }
    if (!cores[set].empty() && !s.gpu.start(set, cores[set].data(), sizes[set].data(), cores[set].size(), true))
      cores[set].clear();
    return !cores[set].empty();
  };
  // Three batches are in flight, each in its own buffer set: while the callbacks consume batch k, a helper thread
  // entropy-decodes (waits for the pool, or uploads the coded bytes of directory-carrying streams) and inverts batch
  // k + 1 on the GPU, and the pool already copies / brotli-decodes batch k + 2.
  int cur = 0;
  size_t end_of[3] = {pos, pos, pos};
  bool have = s.have_delta && gather(cur);
  end_of[cur] = scan;
  std::future<size_t> inflight;
  if (have) inflight = std::async(std::launch::async, [&s, cur, options] { return s.gpu.finish(cur, options); });
  while (have) {
    const int nxt = (cur + 1) % 3;
    const bool have_next = !stream_bad && gather(nxt);     // its pool tasks run while the GPU works on `cur`
    end_of[nxt] = scan;
    const size_t good = inflight.get();
    if (have_next) inflight = std::async(std::launch::async, [&s, nxt, options] { return s.gpu.finish(nxt, options); });
    for (size_t i = 0; i < good; i++) {
      callback(true, s.gpu.outbuf(cur).as<uint16_t>() + i * s.gpu.P, s.gpu.W, s.gpu.H, payload);
      s.id++;
    }
    if (good != cores[cur].size()) {
      if (have_next) inflight.get();      // let the started work drain before the buffers go away
      return fail("decompressing frame failed");
    }
    pos = end_of[cur];
    cur = nxt;
    have = have_next;
  }
  if (stream_bad) return fail(bad_what);

  // keep what was not consumed
  if (s.buffer.empty()) {
    if (pos < size) s.buffer.assign(bytes + pos, bytes + size);
  } else if (pos > 0) {
    s.buffer.erase(s.buffer.begin(), s.buffer.begin() + (ptrdiff_t)pos);
  }
}

// =====================================================================================
// RandomAccessDecoder
// =====================================================================================

struct RandomAccessDecoder::Impl {
  GpuOptions opt;
  mutable std::mutex m;             // DecodeFrame / DecodePreview are const but share GPU staging
  mutable GpuDecoder gpu, preview_gpu;
  size_t xsize = 0, ysize = 0;
  std::vector<uint64_t> offsets;
  const uint8_t* data = nullptr;
  size_t size = 0;

  // Validates the frame chunk at `index`; returns its start and sizes.
  bool locate(size_t index, const uint8_t** chunk, size_t* frame_size, size_t* preview_size) const {
    if (index >= offsets.size()) return FPV_FAIL("invalid frame index");
    const uint64_t off = offsets[index];
    if (off > size || size - off < 9) return FPV_FAIL("out of bounds");
    const uint8_t* p = data + off;
    *frame_size = LoadU32(p);
    if (*frame_size < 9) return FPV_FAIL("frame too small");
    if (size - off < *frame_size) return FPV_FAIL("out of bounds");
    if (p[4] != kChunkFrame) return FPV_FAIL("not a standard frame");
    *preview_size = LoadU32(p + 5);
    if (*preview_size > *frame_size - 9) return FPV_FAIL("preview too large");
    *chunk = p;
    return true;
  }
};

RandomAccessDecoder::RandomAccessDecoder() : RandomAccessDecoder(GpuOptions()) {}
RandomAccessDecoder::RandomAccessDecoder(const GpuOptions& options) : impl_(new Impl) { impl_->opt = options; }
RandomAccessDecoder::~RandomAccessDecoder() = default;

size_t RandomAccessDecoder::xsize() const { return impl_->xsize; }
size_t RandomAccessDecoder::ysize() const { return impl_->ysize; }
size_t RandomAccessDecoder::numframes() const { return impl_->offsets.size(); }

bool RandomAccessDecoder::Init(const uint8_t* data, size_t size) {
  Impl& s = *impl_;
  if (size < 12) return FPV_FAIL("data too small to contain header");
  s.data = data;
  s.size = size;
  s.xsize = LoadU32(data);
  s.ysize = LoadU32(data + 4);
  if (!CheckDims(s.xsize, s.ysize)) return false;

  const size_t chunk = LoadU32(data + 8);
  if (chunk > size - 8) return FPV_FAIL("out of bounds");
  if (chunk < 5) return FPV_FAIL("delta frame too small");
  if (data[12] != kChunkDelta) return FPV_FAIL("must begin with delta frame");
  if (s.gpu.ctx && (s.gpu.W != s.xsize || s.gpu.H != s.ysize)) {
    // re-initialised on a stream of another geometry: the contexts and staging buffers start over
    s.gpu.close();
    s.preview_gpu.close();
  }
  s.offsets.clear();
  if (!s.gpu.ctx && !s.gpu.open(s.opt, s.xsize, s.ysize, 0, false, s.opt.batch)) return false;
  const uint8_t* core = data + 13;
  const size_t core_size = chunk - 5;
  if (s.gpu.decode(&core, &core_size, 1, FPV_DEC_DEFAULT, false) != 1) return FPV_FAIL("failed to decode delta frame");
  if (fpv_set_delta_image(s.gpu.ctx, s.gpu.out.as<uint16_t>()) != FPV_OK) return FPV_FAIL("fpv_set_delta_image");

  // frame index at the end of the file
  if (size < 8 + chunk + 13) return FPV_FAIL("footer missing");
  const uint64_t n = LoadU64(data + size - 8);
  if (n > size / 16) return FPV_FAIL("too many frames");
  const size_t footer = 5 + 8 * (size_t)n + 8;
  if (footer > size) return FPV_FAIL("footer too large");
  const uint8_t* f = data + size - footer;
  if (LoadU32(f) != footer) return FPV_FAIL("footer size mismatch");
  if (f[4] != kChunkIndex) return FPV_FAIL("must end with frame index");
  s.offsets.resize((size_t)n);
  for (size_t i = 0; i < (size_t)n; i++) s.offsets[i] = LoadU64(f + 5 + 8 * i);
  return true;
}

bool RandomAccessDecoder::DecodeFrame(size_t index, uint16_t* frame) const { return DecodeFrames(index, 1, frame); }

bool RandomAccessDecoder::DecodeFrames(size_t first, size_t count, uint16_t* frames) const {
  Impl& s = *impl_;
  std::lock_guard<std::mutex> l(s.m);
  if (!s.gpu.ctx) return FPV_FAIL("decoder not initialised");
  std::vector<const uint8_t*> cores;
  std::vector<size_t> sizes;
  for (size_t done = 0; done < count;) {
    const size_t n = std::min<size_t>(s.gpu.B, count - done);
    cores.clear();
    sizes.clear();
    for (size_t i = 0; i < n; i++) {
      const uint8_t* chunk;
      size_t frame_size, preview_size;
      if (!s.locate(first + done + i, &chunk, &frame_size, &preview_size)) return false;
      cores.push_back(chunk + 9 + preview_size);
      sizes.push_back(frame_size - preview_size - 9);
    }
    if (s.gpu.decode(cores.data(), sizes.data(), n, FPV_DEC_DEFAULT, true) != n) return false;
    memcpy(frames + done * s.gpu.P, s.gpu.out.as<uint16_t>(), n * s.gpu.P * 2);
    done += n;
  }
  return true;
}

bool RandomAccessDecoder::DecodePreview(size_t index, uint8_t* preview) const {
  Impl& s = *impl_;
  std::lock_guard<std::mutex> l(s.m);
  const uint8_t* chunk;
  size_t frame_size, preview_size;
  if (!s.locate(index, &chunk, &frame_size, &preview_size)) return false;
  const size_t pw = s.xsize / 4, ph = s.ysize / 4;
  if (pw == 0 || ph == 0) return FPV_FAIL("invalid image dimensions");
  if (!s.preview_gpu.ctx && !s.preview_gpu.open(s.opt, pw, ph, 0, false, 1)) return false;
  // The preview chunk is a core chunk of a (xsize/4) x (ysize/4) image whose
  // flags never carry USE_DELTA (reference .cc:842, :1061-1064).
  const uint8_t* core = chunk + 9;
  if (s.preview_gpu.decode(&core, &preview_size, 1, FPV_DEC_DEFAULT, false) != 1)
    return FPV_FAIL("failed to decompress preview");
  const uint16_t* img = s.preview_gpu.out.as<uint16_t>();
  for (size_t i = 0; i < pw * ph; i++) preview[i] = (uint8_t)(img[i] >> 8);
  return true;
}

}  // namespace fpvc
