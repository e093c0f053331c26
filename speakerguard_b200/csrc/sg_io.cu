// Caller I/O around the attack loop: what attackMain.py does per batch outside
// `attacker.attack` - dataset/Dataset.py:65-87 (load a wav, crop / zero-pad to wav_length) before it and
// save_audio (attackMain.py:154-166: range heuristic, x 2^15, astype(int16), scipy.io.wavfile.write) after it.
// At > 10^5 utterance-iterations/s the reference's single-threaded per-file Python loops are the wall-clock
// bottleneck of a real run, so:
//   * the float -> PCM16 conversion (per-utterance range test + numpy's cast semantics) is one device kernel,
//     which also halves the device->host bytes;
//   * wav files are written / read by a pool of host threads straight from / into one pinned batch buffer.
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "sg_handle.cuh"

// ---------------------------------------------------------------------------------------------
// device: adv [B, N] float -> pcm [B, N] int16
//   save_audio: if 0.9*max <= 1 and 0.9*min >= -1 the utterance is in [-1,1] scale and is multiplied by 2^15
//   (attackMain.py:155-156); numpy's float32 -> int16 cast truncates toward zero through int32 and keeps the low
//   16 bits (so +1.0 * 32768 becomes -32768, exactly as the reference writes it).
// one CTA per utterance
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ short numpy_cast_i16(float v) {
  const int i32 = (fabsf(v) < 2147483648.f) ? __float2int_rz(v) : (int)0x80000000;   // cvttss2si: NaN / overflow -> INT_MIN
  return (short)(i32 & 0xffff);
}

__global__ void __launch_bounds__(512)
pcm16_quantize_kernel(const float* __restrict__ adv, short* __restrict__ pcm, int N, int* __restrict__ scaled) {
  __shared__ float smax[16], smin[16];
  const float* a = adv + (size_t)blockIdx.x * N;
  short* o = pcm + (size_t)blockIdx.x * N;
  float mx = -INFINITY, mn = INFINITY;
  for (int i = threadIdx.x; i < N; i += blockDim.x) { const float v = a[i]; mx = fmaxf(mx, v); mn = fminf(mn, v); }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
  }
  if ((threadIdx.x & 31) == 0) { smax[threadIdx.x >> 5] = mx; smin[threadIdx.x >> 5] = mn; }
  __syncthreads();
  mx = smax[0]; mn = smin[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mx = fmaxf(mx, smax[w]); mn = fminf(mn, smin[w]); }
  const bool unit_range = (0.9f * mx <= 1.f) && (0.9f * mn >= -1.f);
  const float scale = unit_range ? 32768.f : 1.f;
  if (threadIdx.x == 0 && scaled) scaled[blockIdx.x] = unit_range ? 1 : 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) o[i] = numpy_cast_i16(a[i] * scale);
}

extern "C" int sg_pcm16_quantize(sg_handle* h, const float* adv, int B, int N, int16_t* pcm, int32_t* scaled, sg_stream stream) {
  SG_TRY(sg_check_handle(h, false));
  if (!adv || !pcm || B < 1 || N < 1) { sg_set_error("sg_pcm16_quantize: bad argument"); return SG_EINVAL; }
  h->launches += 1;
  pcm16_quantize_kernel<<<B, 512, 0, (cudaStream_t)stream>>>(adv, (short*)pcm, N, scaled);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

// ---------------------------------------------------------------------------------------------
// host: threaded wav writer / reader (no device work; callable without a GPU)
// ---------------------------------------------------------------------------------------------
static void put_u32(unsigned char* p, uint32_t v) { p[0] = v & 255; p[1] = (v >> 8) & 255; p[2] = (v >> 16) & 255; p[3] = (v >> 24) & 255; }
static void put_u16(unsigned char* p, uint16_t v) { p[0] = v & 255; p[1] = (v >> 8) & 255; }
static uint32_t get_u32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t get_u16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

static int mkdirs_for(const std::string& path) {
  for (size_t i = 1; i < path.size(); ++i)
    if (path[i] == '/') {
      const std::string d = path.substr(0, i);
      if (mkdir(d.c_str(), 0777) != 0 && errno != EEXIST) return -1;
    }
  return 0;
}

// mono PCM16 RIFF file, the 44-byte layout scipy.io.wavfile.write produces for an int16 vector
static int write_wav_pcm16(const char* path, const int16_t* pcm, int n, int fs) {
  if (mkdirs_for(path) != 0) return -1;
  FILE* f = fopen(path, "wb");
  if (!f) return -1;
  unsigned char hd[44];
  const uint32_t data_bytes = (uint32_t)n * 2u;
  memcpy(hd, "RIFF", 4); put_u32(hd + 4, 36 + data_bytes); memcpy(hd + 8, "WAVE", 4);
  memcpy(hd + 12, "fmt ", 4); put_u32(hd + 16, 16); put_u16(hd + 20, 1); put_u16(hd + 22, 1);
  put_u32(hd + 24, (uint32_t)fs); put_u32(hd + 28, (uint32_t)fs * 2u); put_u16(hd + 32, 2); put_u16(hd + 34, 16);
  memcpy(hd + 36, "data", 4); put_u32(hd + 40, data_bytes);
  int ok = fwrite(hd, 1, 44, f) == 44 && fwrite(pcm, 2, (size_t)n, f) == (size_t)n;   // little-endian host
  ok = (fclose(f) == 0) && ok;
  return ok ? 0 : -1;
}

template <typename Fn>
static int run_pool(int n, int nthreads, Fn fn) {
  int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
  nt = nt < 1 ? 1 : (nt > 64 ? 64 : nt);
  if (nt > n) nt = n;
  std::atomic<int> next(0), failed(-1);
  auto work = [&]() {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n) break;
      if (fn(i) != 0) { int exp = -1; failed.compare_exchange_strong(exp, i); }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(work);
  work();
  for (auto& x : th) x.join();
  return failed.load();
}

extern "C" int sg_wav_write_batch(const char* const* paths, const int16_t* pcm, int B, int N, int fs, int nthreads) {
  if (!paths || !pcm || B < 1 || N < 1 || fs < 1) { sg_set_error("sg_wav_write_batch: bad argument"); return SG_EINVAL; }
  const int bad = run_pool(B, nthreads, [&](int i) { return write_wav_pcm16(paths[i], pcm + (size_t)i * N, N, fs); });
  if (bad >= 0) { sg_set_error("sg_wav_write_batch: cannot write '%s' (%s)", paths[bad], strerror(errno)); return SG_ESTATE; }
  return SG_OK;
}

// reads the first channel of a PCM16 RIFF file; returns the number of frames, or -1
static long read_wav_pcm16(const char* path, std::vector<int16_t>& out, int* fs) {
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  unsigned char hd[12];
  if (fread(hd, 1, 12, f) != 12 || memcmp(hd, "RIFF", 4) != 0 || memcmp(hd + 8, "WAVE", 4) != 0) { fclose(f); return -1; }
  int channels = 0, bits = 0, fmt = 0;
  long frames = -1;
  for (;;) {
    unsigned char ch[8];
    if (fread(ch, 1, 8, f) != 8) break;
    const uint32_t sz = get_u32(ch + 4);
    if (memcmp(ch, "fmt ", 4) == 0) {
      unsigned char fm[16];
      if (sz < 16 || fread(fm, 1, 16, f) != 16) break;
      fmt = get_u16(fm); channels = get_u16(fm + 2); *fs = (int)get_u32(fm + 4); bits = get_u16(fm + 14);
      if (sz > 16) fseek(f, (long)(sz - 16 + (sz & 1)), SEEK_CUR);
    } else if (memcmp(ch, "data", 4) == 0) {
      if ((fmt != 1 && fmt != 0xFFFE) || bits != 16 || channels < 1) break;
      const size_t total = sz / 2;
      std::vector<int16_t> raw(total);
      const size_t got = fread(raw.data(), 2, total, f);
      frames = (long)(got / channels);
      out.resize(frames);
      for (long i = 0; i < frames; ++i) out[i] = raw[(size_t)i * channels];
      break;
    } else {
      fseek(f, (long)(sz + (sz & 1)), SEEK_CUR);
    }
  }
  fclose(f);
  return frames;
}

// Dataset.__getitem__ for a batch (dataset/Dataset.py:72-84): out [B, wav_length] float32 (host, e.g. pinned);
// an utterance longer than wav_length is cropped at starts[i] (the caller draws it: np.random.choice(len - wav_length + 1),
// pass -1 to have it centred), a shorter one is zero-padded at the end.  normalize != 0 keeps torchaudio's [-1,1)
// scale (int16 / 2^15), otherwise the int16 range.  lens[i] receives the file's frame count.
extern "C" int sg_wav_read_batch(const char* const* paths, int B, int wav_length, const int64_t* starts, int normalize,
                                 float* out, int32_t* lens, int nthreads) {
  if (!paths || !out || B < 1 || wav_length < 1) { sg_set_error("sg_wav_read_batch: bad argument"); return SG_EINVAL; }
  const int bad = run_pool(B, nthreads, [&](int i) {
    std::vector<int16_t> pcm;
    int fs = 0;
    const long n = read_wav_pcm16(paths[i], pcm, &fs);
    if (n < 0) return -1;
    if (lens) lens[i] = (int32_t)n;
    long start = 0;
    if (n > wav_length) {
      start = (starts && starts[i] >= 0) ? (long)starts[i] : (n - wav_length) / 2;
      if (start > n - wav_length) start = n - wav_length;
    }
    float* o = out + (size_t)i * wav_length;
    const float sc = normalize ? 1.0f / 32768.0f : 1.0f;
    const long keep = n - start < wav_length ? n - start : wav_length;
    for (long k = 0; k < keep; ++k) o[k] = (float)pcm[start + k] * sc;
    for (long k = keep; k < wav_length; ++k) o[k] = 0.f;
    return 0;
  });
  if (bad >= 0) { sg_set_error("sg_wav_read_batch: cannot read '%s' as PCM16 RIFF", paths[bad]); return SG_ESTATE; }
  return SG_OK;
}
