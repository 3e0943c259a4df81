// Reference anchoring: scanSequence / findMaxFreq / the orientation decision of
// getReferenceSlice, reference src/fmindex.h:173-284, for a batch of traces on the GPU.
//
// tracy asks an FM-index (sdsl csa_wt over the upper-cased reference text, sequences joined by '\n',
// src/index.h:104-121, src/fmindex.h:130,160) two things only: count(kmer) and locate(kmer). Both are functions of
// the TEXT, not of the index structure, so the B200 index is the layout HBM likes instead: every text position i with
// the 16 characters that follow it packed into one 64-bit key (4 bits per character, a sequence end pads with 0), all
// positions radix-sorted by key and stored as 16-byte records {key, position}. count(pattern) for |pattern| <= 16 is the
// width of a key range, locate() the position column of that range. A directory over the first d characters (d = 8..14 by
// text size, when they are all A/C/G/T) hands back a bucket of a few records, which is read with independent 16-byte
// loads: a unique k-mer costs one directory access plus one or two 64-byte DRAM granules, no dependent probe chain.
// 16 bytes per text position + 8 * 4^d (52 GB for a 3.1 Gbp genome -- HBM-resident on a 180 GB part).
//
// Per trace one block scans both strands: k-mer -> key range -> hit (position - k) -> a shared-memory hash table of
// hit counts -> findMaxFreq (the smallest most frequent value) -> the reference's orientation rule. Traces that fail
// the unique pass, or whose tables do not fit shared memory, are redone by the global-table kernels (non-unique mode
// pushes every location of k-mers with 0 < occs < 1000, src/fmindex.h:223-229).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace tb {

constexpr int kIdxDepth = 16;                 // characters per key
constexpr int kBucketScan = 8;                // buckets up to this many records are read in one go
constexpr long long kEmpty = LLONG_MIN;       // free hash slot (a hit is position - k, never near INT64_MIN)

// 4-bit alphabet: 0 = end of sequence / padding; the IUPAC letters tracy's texts and consensus strings can hold.
__host__ __device__ __forceinline__ unsigned code4(unsigned char ch) {
  switch (ch) {
    case 'A': return 1; case 'C': return 2; case 'G': return 3; case 'T': return 4; case 'N': return 5;
    case 'R': return 6; case 'Y': return 7; case 'S': return 8; case 'W': return 9; case 'K': return 10;
    case 'M': return 11; case 'B': return 12; case 'D': return 13; case 'H': return 14; case 'V': return 15;
    default: return 0;
  }
}

// ---- index build ------------------------------------------------------------------------------------------------
// key(i) = code4(text[i]) .. code4(text[i+15]), most significant first, everything from the first '\n' / text end on
// padded with 0. Any byte outside the alphabet and '\n' sets *invalid (the build then fails: exactness could not be
// guaranteed for that text).
__global__ void __launch_bounds__(256) index_keys_kernel(const unsigned char* __restrict__ text, long long n,
                                                         unsigned long long* __restrict__ keys, unsigned* __restrict__ pos,
                                                         int* __restrict__ invalid) {
  __shared__ unsigned char tile[256 + kIdxDepth];
  const long long base = (long long)blockIdx.x * 256;
  for (int j = threadIdx.x; j < 256 + kIdxDepth; j += 256) {
    const long long i = base + j;
    const unsigned char ch = i < n ? text[i] : (unsigned char)'\n';
    unsigned c = code4(ch);
    if (c == 0 && ch != '\n' && i < n) atomicExch(invalid, 1);
    tile[j] = (unsigned char)c;
  }
  __syncthreads();
  const long long i = base + threadIdx.x;
  if (i >= n) return;
  unsigned long long key = 0;
  bool live = true;
#pragma unroll
  for (int j = 0; j < kIdxDepth; ++j) {
    const unsigned c = tile[threadIdx.x + j];
    live = live && c != 0;
    key = (key << 4) | (live ? c : 0u);
  }
  keys[i] = key;
  pos[i] = (unsigned)i;
}

__device__ __forceinline__ unsigned long long rec_key(const uint4* __restrict__ rec, long long i) {
  const uint2 k = __ldg(reinterpret_cast<const uint2*>(rec + i));
  return ((unsigned long long)k.y << 32) | k.x;
}
__device__ __forceinline__ long long lower_bound_rec(const uint4* __restrict__ rec, long long lo, long long hi, unsigned long long k) {
  while (lo < hi) {                                       // first index with key >= k
    const long long mid = (lo + hi) >> 1;
    if (rec_key(rec, mid) < k) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// sorted (key, position) columns -> 16-byte records
__global__ void __launch_bounds__(256) index_records_kernel(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ pos,
                                                            long long n, uint4* __restrict__ rec) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = keys[i];
  rec[i] = make_uint4((unsigned)k, (unsigned)(k >> 32), pos[i], 0u);
}

// Directory: for every d-mer over A,C,G,T (2 bits each, first character in the top bits) the sorted range of records
// whose key starts with it.
__global__ void __launch_bounds__(256) index_dir_kernel(const uint4* __restrict__ rec, long long n, int d, uint2* __restrict__ dir) {
  const unsigned b = blockIdx.x * 256u + threadIdx.x;
  if (b >= (1u << (2 * d))) return;
  unsigned long long k = 0;
  for (int j = d - 1; j >= 0; --j) k = (k << 4) | (((b >> (2 * j)) & 3u) + 1u);
  const int pad = 4 * (kIdxDepth - d);
  const unsigned long long klo = k << pad, khi = klo | ((1ull << pad) - 1ull);
  const long long lo = lower_bound_rec(rec, 0, n, klo);
  const long long hi = lower_bound_rec(rec, lo, n, khi + 1ull);   // d <= 14: khi + 1 cannot wrap
  dir[b] = make_uint2((unsigned)lo, (unsigned)hi);
}


// ---- queries ----------------------------------------------------------------------------------------------------

// reference src/fmindex.h:11-26: reversed, upper-cased, A<->T C<->G N->N; any other character leaves the ORIGINAL
// character of that position untouched (`default: break` on a copy made from the forward string).
__device__ __forceinline__ unsigned char strand_char(const char* __restrict__ s, int len, int strand, int i) {
  if (!strand) return (unsigned char)s[i];
  unsigned char c = (unsigned char)s[len - 1 - i];
  if (c >= 'a' && c <= 'z') c -= 32;
  switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'N': return 'N';
    default: return (unsigned char)s[i];
  }
}

// The k-mer starting at k of one strand (reference src/fmindex.h:205-232): returns false when it is skipped (an 'N' in
// the window); otherwise the sorted range [lo, hi) of its occurrences (pattern = substr(k, kmer), shorter at the end)
// and, when the range is not empty, the text position of its first record.
__device__ __forceinline__ bool kmer_range(const KmerIndexView& X, const char* __restrict__ s, int len, int strand, int k, int kmer,
                                           long long* lo_out, long long* hi_out, unsigned* pos_out) {
  const int l = min(kmer, len - k);
  const int d = X.dir_chars;
  unsigned long long key = 0;
  bool findable = true, in_dir = l >= d;
  unsigned dirb = 0;
  for (int j = 0; j < l; ++j) {
    const unsigned char ch = strand_char(s, len, strand, k + j);
    if (ch == 'N') return false;                          // ncount != 0 for this window
    const unsigned c = code4(ch);
    if (c == 0) findable = false;                         // a byte no indexed text contains: count() is 0
    key = (key << 4) | c;
    if (j < d) { if (c >= 1 && c <= 4) dirb = (dirb << 2) | (c - 1); else in_dir = false; }
  }
  *pos_out = 0;
  if (!findable) { *lo_out = *hi_out = 0; return true; }
  const int pad = 4 * (kIdxDepth - l);                    // l >= 1: pad <= 60
  const unsigned long long klo = key << pad, khi = klo | ((1ull << pad) - 1ull);
  long long a = 0, b = X.n;
  if (in_dir) { const uint2 e = __ldg(X.dir + dirb); a = e.x; b = e.y; }
  while (b - a > kBucketScan) {                           // narrow while the probe falls outside the match range
    const long long mid = (a + b) >> 1;
    const unsigned long long km = rec_key(X.rec, mid);
    if (km < klo) a = mid + 1; else if (km > khi) b = mid; else break;
  }
  if (b - a <= kBucketScan) {                             // the whole bucket with independent loads
    uint4 r[kBucketScan];
#pragma unroll
    for (int j = 0; j < kBucketScan; ++j) r[j] = a + j < b ? __ldg(X.rec + a + j) : make_uint4(~0u, ~0u, 0u, 0u);
    int below = 0, upto = 0;
    unsigned p = 0;
#pragma unroll
    for (int j = kBucketScan - 1; j >= 0; --j) {
      const unsigned long long kj = ((unsigned long long)r[j].y << 32) | r[j].x;
      const bool live = a + j < b;
      below += live && kj < klo;
      upto += live && kj <= khi;
      if (live && kj >= klo) p = r[j].z;                  // ends as the position of the first record >= klo
    }
    *lo_out = a + below; *hi_out = a + upto; *pos_out = p;
    return true;
  }
  const long long lo = lower_bound_rec(X.rec, a, b, klo);
  const long long hi = khi == ~0ull ? b : lower_bound_rec(X.rec, lo, b, khi + 1ull);
  *lo_out = lo; *hi_out = hi;
  if (hi > lo) *pos_out = __ldg(&X.rec[lo].z);
  return true;
}

__device__ __forceinline__ unsigned hash_hit(long long v) {
  unsigned long long x = (unsigned long long)v * 0x9E3779B97F4A7C15ull;
  return (unsigned)(x >> 32);
}
template <typename TKey, typename TCnt>
__device__ __forceinline__ void table_insert(TKey* keys, TCnt* cnt, unsigned mask, long long v) {
  unsigned slot = hash_hit(v) & mask;
  while (true) {
    const unsigned long long old = atomicCAS((unsigned long long*)(keys + slot), (unsigned long long)kEmpty, (unsigned long long)v);
    if (old == (unsigned long long)kEmpty || old == (unsigned long long)v) { atomicAdd(cnt + slot, 1u); return; }
    slot = (slot + 1) & mask;
  }
}

// scan range of one strand: k in [tl, kend), reference src/fmindex.h:210 (uint16 k: callers keep len < 65536)
__device__ __forceinline__ void strand_range(int len, int strand, int trim_left, int trim_right, int* k0, int* k1) {
  const int tl = strand ? trim_right : trim_left, tr = strand ? trim_left : trim_right;
  *k0 = tl;
  *k1 = tr <= len ? len - tr : len;                       // size_t wrap of (size - trimRight) leaves only k < size
}

// findMaxFreq over a filled table (reference src/fmindex.h:173-198): highest count, smallest value among those; an
// empty table gives (0, 0). Block-wide; the result is valid in thread 0.
__device__ __forceinline__ void table_mode(const long long* keys, const unsigned* cnt, unsigned size, unsigned* freq, long long* gpos,
                                           unsigned* s_cnt, long long* s_key) {
  unsigned bc = 0; long long bk = 0;
  for (unsigned i = threadIdx.x; i < size; i += blockDim.x) {
    const long long k = keys[i];
    if (k == kEmpty) continue;
    const unsigned c = cnt[i];
    if (c > bc || (c == bc && k < bk)) { bc = c; bk = k; }
  }
  for (int d = 16; d > 0; d >>= 1) {
    const unsigned oc = __shfl_down_sync(0xffffffffu, bc, d);
    const long long ok = __shfl_down_sync(0xffffffffu, bk, d);
    if (oc > bc || (oc == bc && oc > 0 && ok < bk)) { bc = oc; bk = ok; }
  }
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if ((threadIdx.x & 31) == 0) { s_cnt[w] = bc; s_key[w] = bk; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < nw; ++i)
      if (s_cnt[i] > bc || (s_cnt[i] == bc && bc > 0 && s_key[i] < bk)) { bc = s_cnt[i]; bk = s_key[i]; }
    *freq = bc; *gpos = bc ? bk : 0;
  }
  __syncthreads();
}

// getReferenceSlice's orientation rule (reference src/fmindex.h:262-284). Returns 1 forward, 2 reverse, 0 undecided.
__device__ __forceinline__ int orient(unsigned ff, unsigned fr, unsigned min_support) {
  if (ff >= min_support && ff > 2u * fr) return 1;
  if (fr >= min_support && fr > 2u * ff) return 2;
  return 0;
}

// Unique pass, tables in shared memory: one block per trace, both strands one after the other.
// dynamic smem: long long keys[tsize] | unsigned cnt[tsize]
__global__ void __launch_bounds__(256) anchor_unique_kernel(const KmerIndexView X, const AnchorBatch A, unsigned tsize) {
  extern __shared__ __align__(16) unsigned char smem[];
  long long* tkeys = reinterpret_cast<long long*>(smem);
  unsigned* tcnt = reinterpret_cast<unsigned*>(tkeys + tsize);
  __shared__ unsigned s_cnt[8];
  __shared__ long long s_key[8];
  __shared__ unsigned s_freq[2];
  __shared__ long long s_pos[2];
  const int t = blockIdx.x;
  const int len = A.cons_len[t];
  const char* s = A.cons_base + A.cons_off[t];
  for (int strand = 0; strand < 2; ++strand) {
    int k0, k1;
    strand_range(len, strand, A.trim_left, A.trim_right, &k0, &k1);
    if ((unsigned)max(k1 - k0, 0) * 2u > tsize) {         // would not fit: the global-table path redoes this trace
      if (threadIdx.x == 0) A.pass[t] = 3;
      return;
    }
    for (unsigned i = threadIdx.x; i < tsize; i += blockDim.x) { tkeys[i] = kEmpty; tcnt[i] = 0; }
    __syncthreads();
    for (int k = k0 + (int)threadIdx.x; k < k1; k += blockDim.x) {
      long long lo, hi;
      unsigned p0;
      if (!kmer_range(X, s, len, strand, k, A.kmer, &lo, &hi, &p0)) continue;
      if (hi - lo == 1) table_insert(tkeys, tcnt, tsize - 1, (long long)p0 - (long long)k);
    }
    __syncthreads();
    table_mode(tkeys, tcnt, tsize, &s_freq[strand], &s_pos[strand], s_cnt, s_key);
  }
  if (threadIdx.x == 0) {
    const int o = orient(s_freq[0], s_freq[1], (unsigned)A.min_support);
    A.pass[t] = o ? 1 : 2;                                // 1: decided by the unique pass; 2: needs the non-unique pass
    A.anchored[t] = o != 0;
    A.forward[t] = o != 2;
    A.kmersupport[t] = o == 1 ? s_freq[0] : o == 2 ? s_freq[1] : 0;
    A.bestpos[t] = o == 1 ? s_pos[0] : o == 2 ? s_pos[1] : 0;
  }
}

// Global-table path, step 1: how many hits will each (trace, strand) push?
__global__ void __launch_bounds__(256) anchor_count_kernel(const KmerIndexView X, const AnchorBatch A) {
  const int slot = blockIdx.x >> 1, strand = blockIdx.x & 1;
  const int t = A.todo[slot];
  const int len = A.cons_len[t];
  const char* s = A.cons_base + A.cons_off[t];
  int k0, k1;
  strand_range(len, strand, A.trim_left, A.trim_right, &k0, &k1);
  unsigned long long mine = 0;
  for (int k = k0 + (int)threadIdx.x; k < k1; k += blockDim.x) {
    long long lo, hi;
    unsigned p0;
    if (!kmer_range(X, s, len, strand, k, A.kmer, &lo, &hi, &p0)) continue;
    const long long occs = hi - lo;
    if (A.nonunique ? (occs > 0 && occs < 1000) : occs == 1) mine += (unsigned long long)occs;
  }
  for (int d = 16; d > 0; d >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, d);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(A.totals + blockIdx.x, mine);
}

// step 2: push the hits into the (zeroed) global tables
__global__ void __launch_bounds__(256) anchor_fill_kernel(const KmerIndexView X, const AnchorBatch A) {
  const int slot = blockIdx.x >> 1, strand = blockIdx.x & 1;
  const int t = A.todo[slot];
  const int len = A.cons_len[t];
  const char* s = A.cons_base + A.cons_off[t];
  long long* tkeys = A.tab_keys + A.tab_off[blockIdx.x];
  unsigned* tcnt = A.tab_cnt + A.tab_off[blockIdx.x];
  const unsigned mask = A.tab_size[blockIdx.x] - 1;
  int k0, k1;
  strand_range(len, strand, A.trim_left, A.trim_right, &k0, &k1);
  for (int k = k0 + (int)threadIdx.x; k < k1; k += blockDim.x) {
    long long lo, hi;
    unsigned p0;
    if (!kmer_range(X, s, len, strand, k, A.kmer, &lo, &hi, &p0)) continue;
    const long long occs = hi - lo;
    if (!(A.nonunique ? (occs > 0 && occs < 1000) : occs == 1)) continue;
    for (long long j = lo; j < hi; ++j) table_insert(tkeys, tcnt, mask, (long long)__ldg(&X.rec[j].z) - (long long)k);
  }
}

// step 3: findMaxFreq per strand and the orientation rule, one block per trace
__global__ void __launch_bounds__(256) anchor_decide_kernel(const AnchorBatch A) {
  __shared__ unsigned s_cnt[8];
  __shared__ long long s_key[8];
  __shared__ unsigned s_freq[2];
  __shared__ long long s_pos[2];
  const int slot = blockIdx.x;
  const int t = A.todo[slot];
  for (int strand = 0; strand < 2; ++strand) {
    const int b = 2 * slot + strand;
    table_mode(A.tab_keys + A.tab_off[b], A.tab_cnt + A.tab_off[b], A.tab_size[b], &s_freq[strand], &s_pos[strand], s_cnt, s_key);
  }
  if (threadIdx.x == 0) {
    const int o = orient(s_freq[0], s_freq[1], (unsigned)A.min_support);
    if (o || A.nonunique) {
      A.pass[t] = o ? (A.nonunique ? 4 : 1) : 0;          // 4: decided by the non-unique pass; 0: not anchored
      A.anchored[t] = o != 0;
      A.forward[t] = o != 2;
      A.kmersupport[t] = o == 1 ? s_freq[0] : o == 2 ? s_freq[1] : 0;
      A.bestpos[t] = o == 1 ? s_pos[0] : o == 2 ? s_pos[1] : 0;
    } else {
      A.pass[t] = 2;
    }
  }
}

__global__ void fill_empty_kernel(long long* keys, unsigned* cnt, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = kEmpty; cnt[i] = 0; }
}

// ---- launchers (called from capi.cu) ------------------------------------------------------------------------------
cudaError_t index_sort_temp_bytes(long long n, size_t* bytes) {
  return cub::DeviceRadixSort::SortPairs(nullptr, *bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                         (const unsigned*)nullptr, (unsigned*)nullptr, n, 0, 64);
}
// Directory depth for a text of n characters: the smallest d in 8..14 whose 4^d buckets hold <= 8 records on average.
int index_dir_chars(long long n) {
  int d = 8;
  while (d < 14 && (1ll << (2 * d)) * kBucketScan < n) ++d;
  return d;
}
// The build in three steps, so that the caller can release the sort's input columns before the records are allocated
// (peak device memory 28 B per text character instead of 40):
//   index_sort_text : keys/positions from the text into keys_a/pos_a, radix-sorted into keys_b/pos_b
//   index_make_records / index_make_dir : records from the sorted columns, then the directory over the records
cudaError_t index_sort_text(const unsigned char* text, long long n, unsigned long long* keys_a, unsigned* pos_a,
                            unsigned long long* keys_b, unsigned* pos_b, void* temp, size_t temp_bytes, int* invalid, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(invalid, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  index_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(text, n, keys_a, pos_a, invalid);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_a, keys_b, pos_a, pos_b, n, 0, 64, st);
}
cudaError_t index_make_records(const unsigned long long* keys_b, const unsigned* pos_b, long long n, uint4* rec, cudaStream_t st) {
  index_records_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keys_b, pos_b, n, rec);
  return cudaGetLastError();
}
cudaError_t index_make_dir(const uint4* rec, long long n, uint2* dir, int dir_chars, cudaStream_t st) {
  index_dir_kernel<<<(1u << (2 * dir_chars)) / 256, 256, 0, st>>>(rec, n, dir_chars, dir);
  return cudaGetLastError();
}

cudaError_t launch_anchor_unique(const KmerIndexView& X, const AnchorBatch& A, int ntraces, unsigned tsize, cudaStream_t st) {
  const size_t smem = (size_t)tsize * 12;
  cudaError_t e = cudaFuncSetAttribute(anchor_unique_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  anchor_unique_kernel<<<ntraces, 256, smem, st>>>(X, A, tsize);
  return cudaGetLastError();
}
cudaError_t launch_anchor_count(const KmerIndexView& X, const AnchorBatch& A, int ntodo, cudaStream_t st) {
  anchor_count_kernel<<<2 * ntodo, 256, 0, st>>>(X, A);
  return cudaGetLastError();
}
cudaError_t launch_anchor_fill(const KmerIndexView& X, const AnchorBatch& A, int ntodo, long long table_elems, cudaStream_t st) {
  if (table_elems > 0) fill_empty_kernel<<<(unsigned)((table_elems + 255) / 256), 256, 0, st>>>(A.tab_keys, A.tab_cnt, table_elems);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  anchor_fill_kernel<<<2 * ntodo, 256, 0, st>>>(X, A);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  anchor_decide_kernel<<<ntodo, 256, 0, st>>>(A);
  return cudaGetLastError();
}

}  // namespace tb
