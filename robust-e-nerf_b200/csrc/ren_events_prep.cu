// ren_events_prep.cu -- raw event stream -> per-pixel event intervals.
//
// Replaces the two per-event Python loops of the reference's dataset preparation,
//   Event.queue_raw_events              robust_e_nerf/data/datasets.py:189-284
//   Event.extract_max_refractory_period robust_e_nerf/data/datasets.py:132-187
// Both walk the time-ordered stream with a 2-deep sliding window PER PIXEL: event i is paired with
// the previous event at the same pixel; it is kept iff that predecessor exists and has a different
// timestamp (start_ts = predecessor timestamp, end_ts = own timestamp, num_pos = own polarity,
// num_neg = 1 - num_pos); the maximum refractory period is the smallest positive interval between
// consecutive events of any pixel.  Integer work, bit-exact.
//
// Data-parallel form: a STABLE radix sort of (pixel id, event index) keeps the time order inside each
// pixel, so "previous event at the same pixel" is the left neighbour in sorted order; validity flags
// are compacted in original event order with an exclusive scan.  HBM-bound: ~60 B per event.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "ren_common.cuh"

struct PrepArgs {
  long long n;
  int width, height;
  const unsigned short* pos;     // [n][2] (x, y)
  const long long* ts;           // [n]
  const unsigned char* pol;      // [n] 0/1
  unsigned int* keys;            // [n]
  unsigned int* vals;            // [n]
  const unsigned int* keys_s;    // sorted
  const unsigned int* vals_s;
  unsigned int* flags;           // [n] valid (0/1), original order
  const unsigned int* offs;      // [n] exclusive scan of flags
  long long* start_tmp;          // [n] predecessor timestamp, original order
  long long* min_interval;       // [1] (initialised to LLONG_MAX)
  long long* out_position;       // [n_valid][2] i64
  long long* out_start;
  long long* out_end;
  long long* out_num_pos;
  long long* out_num_neg;
  long long* n_valid;            // [1]
};

extern "C" __global__ void ren_prep_init(long long* min_interval, long long* n_valid) {
  *min_interval = 0x7fffffffffffffffLL;
  *n_valid = 0;
}

extern "C" __global__ void ren_prep_keys(const PrepArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const unsigned int x = a.pos[2 * i], y = a.pos[2 * i + 1];
  a.keys[i] = y * (unsigned int)a.width + x;
  a.vals[i] = (unsigned int)i;
}

extern "C" __global__ void ren_prep_pair(const PrepArgs a) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long best = 0x7fffffffffffffffLL;
  if (j < a.n) {
    const unsigned int i = a.vals_s[j];
    unsigned int valid = 0;
    long long prev = 0;
    if (j > 0 && a.keys_s[j - 1] == a.keys_s[j]) {
      prev = a.ts[a.vals_s[j - 1]];
      const long long cur = a.ts[i];
      valid = prev != cur;                              // datasets.py:247-254
      if (cur != prev) best = cur - prev;               // datasets.py:165-186 (duplicates are skipped)
    }
    a.flags[i] = valid;
    a.start_tmp[i] = prev;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long t = __shfl_xor_sync(0xffffffffu, best, o);
    best = t < best ? t : best;
  }
  if ((threadIdx.x & 31) == 0 && best != 0x7fffffffffffffffLL) atomicMin(a.min_interval, best);
}

extern "C" __global__ void ren_prep_compact(const PrepArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  if (i == a.n - 1) a.n_valid[0] = (long long)a.offs[i] + a.flags[i];
  if (!a.flags[i]) return;
  const long long o = a.offs[i];
  a.out_position[2 * o] = a.pos[2 * i];
  a.out_position[2 * o + 1] = a.pos[2 * i + 1];
  a.out_start[o] = a.start_tmp[i];
  a.out_end[o] = a.ts[i];
  const long long p = a.pol[i] ? 1 : 0;
  a.out_num_pos[o] = p;                                 // sum(window) - window[0] == own polarity
  a.out_num_neg[o] = 1 - p;                             // (maxlen - 1) - num_pos
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct PrepLayout {
  size_t keys, vals, keys_s, vals_s, flags, offs, start_tmp, min_interval, cub, total, cub_bytes;
};
static int prep_layout(long long n, int n_bits, PrepLayout* L) {
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned int*)nullptr, (unsigned int*)nullptr,
                                  (const unsigned int*)nullptr, (unsigned int*)nullptr, (int)n, 0, n_bits);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const unsigned int*)nullptr, (unsigned int*)nullptr, (int)n);
  size_t off = 0;
  L->keys = off; off += al256(4 * (size_t)n);
  L->vals = off; off += al256(4 * (size_t)n);
  L->keys_s = off; off += al256(4 * (size_t)n);
  L->vals_s = off; off += al256(4 * (size_t)n);
  L->flags = off; off += al256(4 * (size_t)n);
  L->offs = off; off += al256(4 * (size_t)n);
  L->start_tmp = off; off += al256(8 * (size_t)n);
  L->min_interval = off; off += 256;
  L->cub = off;
  L->cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
  off += al256(L->cub_bytes);
  L->total = off + 256;
  return 0;
}
static int key_bits(int width, int height) {
  long long cells = (long long)width * height;
  int b = 1;
  while ((1LL << b) < cells) ++b;
  return b;
}

extern "C" int ren_event_queue_workspace_bytes(int64_t n_events, int32_t img_width, int32_t img_height,
                                               size_t* bytes) {
  if (n_events < 0 || n_events >= (1LL << 31) || img_width < 1 || img_height < 1 ||
      (long long)img_width * img_height >= (1LL << 32) || !bytes) {
    ren_set_error("ren_event_queue_workspace_bytes: bad arguments");
    return REN_E_INVALID;
  }
  PrepLayout L;
  prep_layout(n_events > 0 ? n_events : 1, key_bits(img_width, img_height), &L);
  *bytes = L.total;
  return REN_OK;
}

extern "C" int ren_event_queue(int64_t n_events, int32_t img_width, int32_t img_height,
                               const uint16_t* position, const int64_t* timestamp, const uint8_t* polarity,
                               int64_t* out_position, int64_t* out_start_ts, int64_t* out_end_ts,
                               int64_t* out_num_pos, int64_t* out_num_neg, int64_t* out_n_valid,
                               int64_t* out_min_interval, void* workspace, size_t workspace_bytes, void* stream) {
  if (n_events < 0 || n_events >= (1LL << 31) || img_width < 1 || img_height < 1 ||
      (long long)img_width * img_height >= (1LL << 32)) {
    ren_set_error("ren_event_queue: bad sizes");
    return REN_E_INVALID;
  }
  if (!out_n_valid || !out_min_interval) { ren_set_error("ren_event_queue: null output"); return REN_E_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  ren_prep_init<<<1, 1, 0, st>>>((long long*)out_min_interval, (long long*)out_n_valid);
  if (n_events == 0) return REN_OK;
  if (!position || !timestamp || !polarity || !out_position || !out_start_ts || !out_end_ts || !out_num_pos ||
      !out_num_neg || !workspace) {
    ren_set_error("ren_event_queue: null pointer argument");
    return REN_E_INVALID;
  }
  const int n_bits = key_bits(img_width, img_height);
  PrepLayout L;
  prep_layout(n_events, n_bits, &L);
  char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  if ((size_t)(ws - (char*)workspace) + L.total - 256 > workspace_bytes) {
    ren_set_error("ren_event_queue: workspace too small (%zu needed)", L.total);
    return REN_E_WORKSPACE;
  }
  PrepArgs a;
  a.n = n_events; a.width = img_width; a.height = img_height;
  a.pos = position; a.ts = (const long long*)timestamp; a.pol = polarity;
  a.keys = (unsigned int*)(ws + L.keys); a.vals = (unsigned int*)(ws + L.vals);
  a.keys_s = (unsigned int*)(ws + L.keys_s); a.vals_s = (unsigned int*)(ws + L.vals_s);
  a.flags = (unsigned int*)(ws + L.flags); a.offs = (unsigned int*)(ws + L.offs);
  a.start_tmp = (long long*)(ws + L.start_tmp); a.min_interval = (long long*)out_min_interval;
  a.out_position = (long long*)out_position; a.out_start = (long long*)out_start_ts;
  a.out_end = (long long*)out_end_ts; a.out_num_pos = (long long*)out_num_pos;
  a.out_num_neg = (long long*)out_num_neg; a.n_valid = (long long*)out_n_valid;
  const unsigned blocks = (unsigned)((n_events + 255) / 256);
  ren_prof_begin(REN_K_EVENT, st);
  ren_prep_keys<<<blocks, 256, 0, st>>>(a);
  ren_prof_end(REN_K_EVENT, st);
  size_t cub_bytes = L.cub_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(ws + L.cub, cub_bytes, a.keys, (unsigned int*)a.keys_s, a.vals,
                                                  (unsigned int*)a.vals_s, (int)n_events, 0, n_bits, st);
  if (e != cudaSuccess) { ren_set_error("ren_event_queue sort: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  ren_prof_begin(REN_K_EVENT, st);
  ren_prep_pair<<<blocks, 256, 0, st>>>(a);
  ren_prof_end(REN_K_EVENT, st);
  cub_bytes = L.cub_bytes;
  e = cub::DeviceScan::ExclusiveSum(ws + L.cub, cub_bytes, a.flags, (unsigned int*)a.offs, (int)n_events, st);
  if (e != cudaSuccess) { ren_set_error("ren_event_queue scan: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  ren_prof_begin(REN_K_EVENT, st);
  ren_prep_compact<<<blocks, 256, 0, st>>>(a);
  ren_prof_end(REN_K_EVENT, st);
  e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("ren_event_queue launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}
