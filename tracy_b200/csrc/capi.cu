// C ABI of tracy_b200 (include/tracy_b200.h): context, scratch sizing, host<->device staging, launches.
// Host code only; the kernels live in gotoh_general.cu / gotoh_packed.cu / sweep.cu.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tracy_b200.h"
#include "common.cuh"

namespace tb {
cudaError_t launch_gotoh_general(int mode, bool traceback, const GotohBatch& B, int blocks, cudaStream_t stream);
cudaError_t gotoh_general_blocks_per_sm(int mode, bool traceback, int* out);
int gotoh_general_warps_per_block();
cudaError_t launch_gotoh_packed(int tbmode, int classes, const GotohBatch& B, int blocks, cudaStream_t stream);
cudaError_t gotoh_packed_blocks_per_sm(int tbmode, int classes, int* out);
int gotoh_packed_warps_per_block();
bool gotoh_packed_eligible(int maxm, int maxn, int match, int mismatch, int go, int ge);
unsigned long long gotoh_packed_ptr_words(int m, int n);
cudaError_t launch_gotoh_pp(int nch, bool traceback, bool harr, const GotohBatch& B, const PPWork& W, int blocks, cudaStream_t stream);
cudaError_t gotoh_pp_blocks_per_sm(int nch, bool traceback, bool harr, int* out);
int gotoh_pp_warps_per_block();


cudaError_t launch_sweep(const SweepBatch& S, int ntraces, bool grid, long long max_tasks, cudaStream_t stream);
void scan_trace_file(const uint8_t* buf, int64_t n, TraceDesc* d);
cudaError_t launch_trace_unpack(const TraceUnpack& U, int nfiles, cudaStream_t st);
cudaError_t launch_allelic_fraction(const FractionBatch& F, int ntraces, int maxD, cudaStream_t st);
cudaError_t index_sort_temp_bytes(long long n, size_t* bytes);
cudaError_t index_sort_text(const unsigned char* text, long long n, unsigned long long* keys_a, unsigned* pos_a, unsigned long long* keys_b,
                            unsigned* pos_b, void* temp, size_t temp_bytes, int* invalid, cudaStream_t st);
cudaError_t index_make_records(const unsigned long long* keys_b, const unsigned* pos_b, long long n, uint4* rec, cudaStream_t st);
cudaError_t index_make_dir(const uint4* rec, long long n, uint2* dir, int dir_chars, cudaStream_t st);
int index_dir_chars(long long n);
cudaError_t launch_anchor_unique(const KmerIndexView& X, const AnchorBatch& A, int ntraces, unsigned tsize, cudaStream_t st);
cudaError_t launch_anchor_count(const KmerIndexView& X, const AnchorBatch& A, int ntodo, cudaStream_t st);
cudaError_t launch_anchor_fill(const KmerIndexView& X, const AnchorBatch& A, int ntodo, long long table_elems, cudaStream_t st);
cudaError_t launch_create_profile(const ProfileBatch& P, int ntraces, cudaStream_t stream);
cudaError_t launch_basecall(const BasecallBatch& P, int ntraces, cudaStream_t stream);
cudaError_t launch_revcomp_profile(const float* in_base, const int64_t* in_off, const int32_t* len, float* out_base, const int64_t* out_off,
                                   int n, cudaStream_t stream);
}  // namespace tb

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    // grow geometrically: a series of calls with growing shapes (the levels of a guide tree: each merge about twice the
    // last) would otherwise free and allocate hundreds of MB per call -- 145 ms for 0.5 GB on the pool's boxes
    const size_t old = cap;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = std::max(bytes + bytes / 8 + 256, std::min<size_t>(2 * old, (size_t)4 << 30));
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&p, bytes); want = bytes; }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T* as() const { return static_cast<T*>(p); }
};
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    const size_t old = cap;
    if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
    size_t want = std::max(bytes + 256, std::min<size_t>(2 * old, (size_t)1 << 30));
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); want = bytes + 256; e = cudaHostAlloc(&p, want, cudaHostAllocDefault); }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Plan {
  bool use_pp = false, pp_harr = false;    // profile x profile: the fp32x2 kernel (gotoh_pp.cu) ahead of the general one
  int blocks_pp = 0, blocks_pp5 = 0;
  bool use_packed = false;
  int tbmode = 0;                          // packed kernel traceback mode: 0 none, 1 flags, 2 checkpoints (default)
  int blocks_packed = 0, blocks_packed5 = 0, blocks_general = 0;   // packed: 4-class (ACGT) and 5-class (ACGTN) instantiations
  unsigned slots = 0;                      // warp slots that own scratch
  unsigned long long ptr_words = 0, rowbuf_elems = 0, ops_bytes = 0;
};

struct Lane {   // one in-flight chunk of a TB_MEM_HOST batch
  cudaStream_t stream = nullptr;
  cudaEvent_t c0 = nullptr, k0 = nullptr, k1 = nullptr, k2 = nullptr;   // c0: start of a device-mode call; k*: kernel brackets
  cudaEvent_t h0 = nullptr, d1 = nullptr;                               // host-mode chunk: before its H2D, after its D2H (TRACY_B200_TRACE)
  cudaEvent_t in_done = nullptr;                                        // host-mode chunk: its inputs have arrived (the next chunk's copy waits for it)
  DevBuf a, b, meta_d, scores, ops, ops_len, status, counter;
  DevBuf ptr, rowbuf, opsrev;
  DevBuf row0, row1, opk;                  // host-mode outputs made by post_ops.cu: gapped rows, 2-bit packed ops
  DevBuf pp_units, pp_small, pp_big, pp_rowbuf, pp_ptr, pp_flags;   // big-pair work list and scratch of the profile x profile kernel
  PinBuf pp_stage;
  DevBuf gate; PinBuf gate_host;           // streamed host batches: gate words of the packed kernel, and their host side
  tb::PPWork ppw{};
  PinBuf meta, cnt;
  bool timed = false, timed2 = false;
  cudaEvent_t kend = nullptr;        // the event after the last kernel enqueued for the lane's current call/chunk
  tb::GotohBatch view{};             // what enqueue_gotoh launched on (finish_gotoh's stage 2 reuses it)
  Plan plan; int mode = 0; bool traceback = false;
  long chunk = -1;   // index of the host-mode chunk in flight on this lane (-1: none)
  size_t p0 = 0;     // its first pair
};
constexpr int kLanes = 4;   // lanes that exist; a TB_MEM_HOST batch keeps `nlanes` (default 3) chunks in flight: copying in, computing, copying out

}  // namespace

struct tb_ctx {
  int device = 0;
  int sms = 0;
  size_t scratch_limit = 0;
  std::string err;
  Lane lanes[kLanes];
  cudaEvent_t t0 = nullptr;   // start of a host-mode call (timeline origin for TRACY_B200_TRACE)
  uint64_t launches = 0, h2d = 0, d2h = 0;
  float last_fast_ms = 0, last_general_ms = 0, last_sweep_ms = 0, last_call_ms = 0, last_anchor_ms = 0, last_fraction_ms = 0;
  uint64_t last_packed_pairs = 0;
  std::vector<int32_t> tmp_len1, tmp_len2;
  int occ_general[3][2] = {{-1, -1}, {-1, -1}, {-1, -1}};   // blocks per SM, general kernel [mode][traceback]
  int occ_packed[3][2] = {{-1, -1}, {-1, -1}, {-1, -1}};    // packed kernel [tbmode][4 / 5 classes]
  int occ_pp[2][2][2] = {{{-1, -1}, {-1, -1}}, {{-1, -1}, {-1, -1}}};   // profile x profile kernel [4 / 5 channels][traceback][register-array variant]
  uint64_t last_pp_pairs = 0, last_big_pairs = 0;
  size_t free_at_first_plan = 0;
  int stream_backoff = 0;     // host batches: calls left before the streamed form is tried again (its inputs arrived no faster than they were used)
};

namespace {

int fail(tb_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
int cuda_fail(tb_ctx* c, cudaError_t e, const char* what) {
  cudaGetLastError();
  int code = (e == cudaErrorMemoryAllocation) ? TB_ERR_NOMEM : TB_ERR_CUDA;
  return fail(c, code, std::string(what) + ": " + cudaGetErrorString(e));
}
#define TB_CUDA(ctx, call)                                   \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
  } while (0)

struct Shape {            // maxima over (a slice of) a batch
  int maxm = 0, maxn = 0, maxsum = 0;
  unsigned long long gen_words = 0, packed_words = 0;
};
inline unsigned long long general_ptr_words(int m, int n) {
  if (m <= 0 || n <= 0) return 0;
  unsigned long long nb = (unsigned long long)(m + 511) / 512;
  return nb * (unsigned long long)(n + 31) * 32ull;
}
// big (optional): pairs that the profile x profile kernel spreads over many warps with per-pair scratch; they do not size
// the per-warp-slot pointer scratch
void accumulate(Shape& s, const int32_t* l1, const int32_t* l2, size_t n, bool packed, const uint8_t* big = nullptr) {
  for (size_t i = 0; i < n; ++i) {
    const int m = l1[i], k = l2[i];
    s.maxm = std::max(s.maxm, m);
    s.maxn = std::max(s.maxn, k);
    s.maxsum = std::max(s.maxsum, m + k);
    if (!big || !big[i]) s.gen_words = std::max(s.gen_words, general_ptr_words(m, k));
    if (packed) s.packed_words = std::max(s.packed_words, tb::gotoh_packed_ptr_words(m, k));
  }
}


// Size the persistent grids and the per-slot scratch for a launch over `npairs` pairs with maxima `sh`.
// pp_tickets: work units of the profile x profile kernel (whole small pairs + bands of big pairs); 0 = npairs.
int make_plan(tb_ctx* ctx, int mode, bool traceback, const Shape& sh, size_t npairs, tb_score sc, Plan* out, size_t pp_tickets = 0) {
  Plan p;
  // occupancy answers never change for a context: ask the runtime once per instantiation (a small call should not pay
  // five occupancy queries and a cudaMemGetInfo every time)
  int& bps_g = ctx->occ_general[mode][traceback ? 1 : 0];
  if (bps_g < 0) TB_CUDA(ctx, tb::gotoh_general_blocks_per_sm(mode, traceback, &bps_g));
  if (bps_g < 1) return fail(ctx, TB_ERR_CUDA, "general kernel does not fit on an SM");
  const int wpb_g = tb::gotoh_general_warps_per_block();
  p.use_packed = (mode == tb::kModePS || mode == tb::kModeSS) &&   // string x string: a1's characters act as one-hot profile columns
                 tb::gotoh_packed_eligible(sh.maxm, sh.maxn, sc.match, sc.mismatch, sc.gap_open, sc.gap_extend) &&
                 getenv("TRACY_B200_NO_PACKED") == nullptr;
  unsigned long long warps_pp = 0;
  const int wpb_pp = tb::gotoh_pp_warps_per_block();
  // (the pp kernel carries the vertical state in a form that needs gap open <= 0; anything else goes to the general kernel)
  if (mode == tb::kModePP && sc.gap_open <= 0 && getenv("TRACY_B200_NO_PPFAST") == nullptr) {
    // free-row costs in register arrays (2 blocks / SM; the default: with the screened score the cell loop is bound by the
    // integer pipe, and the selects were a third of its instructions) or, "sel", as selects (3 blocks / SM)
    const char* v = getenv("TRACY_B200_PP_VARIANT");
    p.pp_harr = !(v && std::strcmp(v, "sel") == 0);
    int& o4 = ctx->occ_pp[0][traceback ? 1 : 0][p.pp_harr ? 1 : 0];
    int& o5 = ctx->occ_pp[1][traceback ? 1 : 0][p.pp_harr ? 1 : 0];
    if (o4 < 0) TB_CUDA(ctx, tb::gotoh_pp_blocks_per_sm(4, traceback, p.pp_harr, &o4));
    if (o5 < 0) TB_CUDA(ctx, tb::gotoh_pp_blocks_per_sm(5, traceback, p.pp_harr, &o5));
    if (o4 >= 1 && o5 >= 1) {
      p.use_pp = true;
      p.blocks_pp = ctx->sms * o4; p.blocks_pp5 = ctx->sms * o5;
      warps_pp = (unsigned long long)std::max(p.blocks_pp, p.blocks_pp5) * wpb_pp;
    }
  }
  int bps_p = 0, bps_p5 = 0, wpb_p = 1;
  if (p.use_packed) {
    const char* mode_env = getenv("TRACY_B200_TB_MODE");   // "flags" selects the pointer-flag fill (kept for comparison / profiling)
    p.tbmode = !traceback ? 0 : (mode_env && std::strcmp(mode_env, "flags") == 0) ? 1 : 2;
    int& c4 = ctx->occ_packed[p.tbmode][0];
    int& c5 = ctx->occ_packed[p.tbmode][1];
    if (c4 < 0) TB_CUDA(ctx, tb::gotoh_packed_blocks_per_sm(p.tbmode, 4, &c4));
    if (c5 < 0) TB_CUDA(ctx, tb::gotoh_packed_blocks_per_sm(p.tbmode, 5, &c5));
    bps_p = c4; bps_p5 = c5;
    wpb_p = tb::gotoh_packed_warps_per_block();
    if (bps_p < 1 || bps_p5 < 1) p.use_packed = false;
  }
  unsigned long long warps_g = (unsigned long long)ctx->sms * bps_g * wpb_g;
  unsigned long long warps_p = p.use_packed ? (unsigned long long)ctx->sms * bps_p * wpb_p : 0;
  unsigned long long warps_p5 = p.use_packed ? (unsigned long long)ctx->sms * bps_p5 * wpb_p : 0;
  warps_g = std::min<unsigned long long>(warps_g, std::max<size_t>(npairs, 1));
  warps_p = std::min<unsigned long long>(warps_p, std::max<size_t>(npairs, 1));
  warps_p5 = std::min<unsigned long long>(warps_p5, std::max<size_t>(npairs, 1));

  p.ptr_words = traceback ? std::max(sh.gen_words, sh.packed_words) : 0;
  p.rowbuf_elems = 2ull * (unsigned long long)(sh.maxn + 1);
  p.ops_bytes = traceback ? (((unsigned long long)sh.maxsum + 63) & ~63ull) : 0;
  const unsigned long long per_slot = p.ptr_words * 8 + p.rowbuf_elems * 8 + p.ops_bytes;
  size_t limit = ctx->scratch_limit;
  if (limit == 0) {
    if (ctx->free_at_first_plan == 0) {                     // free HBM when the context first planned (its own scratch not yet taken)
      size_t fr = 0, tot = 0;
      TB_CUDA(ctx, cudaMemGetInfo(&fr, &tot));
      ctx->free_at_first_plan = fr;
    }
    limit = ctx->free_at_first_plan / 3;
  }
  limit /= 3;   // up to three lanes hold scratch at a time by default
  unsigned long long max_slots = per_slot ? std::max<unsigned long long>(limit / per_slot, 1) : (1ull << 30);
  warps_g = std::min(warps_g, max_slots);
  warps_p = std::min(warps_p, max_slots);
  warps_p5 = std::min(warps_p5, max_slots);
  if (p.use_pp) {   // cut to the number of work units, not pairs: a big pair is many units (bands), each taken by a warp
    warps_pp = std::min<unsigned long long>(warps_pp, std::max<size_t>(pp_tickets ? pp_tickets : npairs, 1));
    warps_pp = std::min(warps_pp, max_slots);
    const int cap = (int)((warps_pp + wpb_pp - 1) / wpb_pp);
    p.blocks_pp = std::min(p.blocks_pp, cap); p.blocks_pp5 = std::min(p.blocks_pp5, cap);
  }
  p.blocks_general = (int)((warps_g + wpb_g - 1) / wpb_g);
  p.blocks_packed = p.use_packed ? (int)((warps_p + wpb_p - 1) / wpb_p) : 0;
  p.blocks_packed5 = p.use_packed ? (int)((warps_p5 + wpb_p - 1) / wpb_p) : 0;
  p.slots = (unsigned)std::max<unsigned long long>((unsigned long long)p.blocks_general * wpb_g,
                                                   (unsigned long long)std::max(p.blocks_packed, p.blocks_packed5) * wpb_p);
  if (p.use_pp) p.slots = std::max<unsigned>(p.slots, (unsigned)std::max(p.blocks_pp, p.blocks_pp5) * wpb_pp);
  *out = p;
  return TB_OK;
}

int reserve_scratch(tb_ctx* ctx, Lane& L, const Plan& p) {
  TB_CUDA(ctx, L.ptr.reserve((size_t)(p.ptr_words * 8ull * p.slots)));
  TB_CUDA(ctx, L.rowbuf.reserve((size_t)(p.rowbuf_elems * 8ull * p.slots)));
  TB_CUDA(ctx, L.opsrev.reserve((size_t)(p.ops_bytes * p.slots)));
  TB_CUDA(ctx, L.counter.reserve(64));
  return TB_OK;
}

// ---- profile x profile: big pairs (one pair over many warps, gotoh_pp.cu) ---------------------------------------------
constexpr int kPPBandRows = 512;
inline int pp_bands(int m) { return (m + kPPBandRows - 1) / kPPBandRows; }

// big[i] = 1: pair i is cut into (pair, band) units. A pair on one warp takes bands x columns steps whatever else the GPU
// does, so when the call cannot fill the machine with whole pairs, every pair of three bands or more is spread out; in a
// large batch only outliers are (sixteen bands or more). The per-pair scratch (bottom row per band, pointer words) of the
// pairs chosen stays within `budget` bytes, largest pairs first. Returns the number of work units (tickets).
size_t classify_pp(const tb_ctx* ctx, const int32_t* l1, const int32_t* l2, size_t np, bool traceback, size_t budget, std::vector<uint8_t>& big) {
  big.assign(np, 0);
  const size_t machine = (size_t)ctx->sms * 12;
  const int min_bands = np >= 2 * machine ? 16 : 3;
  std::vector<size_t> cand;
  for (size_t i = 0; i < np; ++i)
    if (l1[i] > 0 && l2[i] >= 256 && pp_bands(l1[i]) >= min_bands) cand.push_back(i);
  std::sort(cand.begin(), cand.end(), [&](size_t x, size_t y) {
    const long long cx = (long long)l1[x] * l2[x], cy = (long long)l1[y] * l2[y];
    return cx != cy ? cx > cy : x < y;
  });
  size_t used = 0, units = np;
  for (size_t i : cand) {
    const size_t nb = (size_t)pp_bands(l1[i]);
    const size_t need = nb * ((size_t)l2[i] + 1) * 8 + (traceback ? nb * ((size_t)l2[i] + 31) * 256 : 0) + nb * 4;
    if (used + need > budget) continue;
    used += need; big[i] = 1; units += nb - 1;
  }
  return units;
}

// Work list and per-pair scratch of the big pairs of one launch (pairs p0 .. p0+cn-1 of the call, indices relative to p0).
int build_pp_work(tb_ctx* ctx, Lane& L, const int32_t* l1, const int32_t* l2, const uint8_t* big, size_t cn, bool traceback) {
  tb::PPWork W{};
  W.one = 1.0f; W.nsmall = (int)cn;
  { const char* v = std::getenv("TRACY_B200_PP_SCREEN"); W.screen = (v && v[0] == '0') ? 0 : 1; }
  size_t nbig = 0, nunits = 0;
  if (big) for (size_t i = 0; i < cn; ++i) if (big[i]) { ++nbig; nunits += (size_t)pp_bands(l1[i]); }
  if (nbig == 0) { L.ppw = W; return TB_OK; }
  const size_t nsmall = cn - nbig;
  const size_t o_units = 0, o_small = o_units + nunits * sizeof(tb::PPUnit), o_big = (o_small + nsmall * 4 + 15) & ~(size_t)15, total = o_big + nbig * sizeof(tb::PPBig);
  TB_CUDA(ctx, L.pp_stage.reserve(total));
  char* h = static_cast<char*>(L.pp_stage.p);
  tb::PPUnit* hu = reinterpret_cast<tb::PPUnit*>(h + o_units);
  int32_t* hs = reinterpret_cast<int32_t*>(h + o_small);
  tb::PPBig* hb = reinterpret_cast<tb::PPBig*>(h + o_big);
  std::vector<size_t> order;
  for (size_t i = 0; i < cn; ++i) if (big[i]) order.push_back(i);
  std::sort(order.begin(), order.end(), [&](size_t x, size_t y) {            // largest first: its pipeline is the longest
    const long long cx = (long long)l1[x] * l2[x], cy = (long long)l1[y] * l2[y];
    return cx != cy ? cx > cy : x < y;
  });
  long long rows = 0, words = 0; int flags = 0; size_t u = 0;
  for (size_t k = 0; k < order.size(); ++k) {
    const size_t i = order[k];
    const int nb = pp_bands(l1[i]);
    hb[k].rowbuf_off = rows; hb[k].ptr_off = words; hb[k].flag_off = flags; hb[k].nb = nb;
    rows += (long long)nb * (l2[i] + 1);
    if (traceback) words += (long long)nb * (l2[i] + 31) * 32;
    flags += nb;
    for (int b = 0; b < nb; ++b) hu[u++] = tb::PPUnit{(int32_t)i, b, (int32_t)k, 0};
  }
  size_t ns = 0;
  for (size_t i = 0; i < cn; ++i) if (!big[i]) hs[ns++] = (int32_t)i;
  TB_CUDA(ctx, L.pp_units.reserve(total));
  TB_CUDA(ctx, L.pp_rowbuf.reserve((size_t)rows * 8 + 64));
  TB_CUDA(ctx, L.pp_ptr.reserve((size_t)words * 8 + 64));
  TB_CUDA(ctx, L.pp_flags.reserve((size_t)flags * 4 + 64));
  TB_CUDA(ctx, cudaMemcpyAsync(L.pp_units.p, h, total, cudaMemcpyHostToDevice, L.stream));
  TB_CUDA(ctx, cudaMemsetAsync(L.pp_flags.p, 0, (size_t)flags * 4, L.stream));
  ctx->h2d += total;
  char* d = L.pp_units.as<char>();
  W.units = reinterpret_cast<const tb::PPUnit*>(d + o_units); W.nunits = (int)nunits;
  W.small_ids = reinterpret_cast<const int32_t*>(d + o_small); W.nsmall = (int)nsmall;
  W.big = reinterpret_cast<const tb::PPBig*>(d + o_big);
  W.big_rowbuf = L.pp_rowbuf.as<int2>(); W.big_ptr = L.pp_ptr.as<unsigned long long>(); W.big_flags = L.pp_flags.as<int>();
  L.ppw = W;
  ctx->last_big_pairs += nbig;
  return TB_OK;
}

// Enqueue the DP kernels for one device-resident batch view on lane L's stream.
// Stage 1 is ONE kernel: the 4-class packed kernel when the batch is eligible for it (it completes every pair whose
// window is pure ACGT and whose score range fits 16 bits), the general kernel otherwise. Whether stage 2 is needed --
// the 5-class packed kernel and the general kernel over the pairs stage 1 left -- is read from the completion counter
// after the stream has drained: persistent grids fill every SM, so empty follow-up launches on the same stream would
// have to wait for the NEXT chunk's kernel (another lane) to retire and would serialise the lanes.
int enqueue_gotoh(tb_ctx* ctx, Lane& L, int mode, bool traceback, tb::GotohBatch B, const Plan& p) {
  B.ptr_scratch = L.ptr.as<unsigned long long>(); B.ptr_slot_words = p.ptr_words;
  B.rowbuf = L.rowbuf.as<int2>(); B.rowbuf_slot = p.rowbuf_elems;
  B.ops_scratch = L.opsrev.as<uint8_t>(); B.ops_slot = p.ops_bytes;
  unsigned int* counters = L.counter.as<unsigned int>();
  TB_CUDA(ctx, cudaEventRecord(L.c0, L.stream));
  TB_CUDA(ctx, cudaMemsetAsync(counters, 0, 64, L.stream));
  TB_CUDA(ctx, cudaMemsetAsync(B.status, 0, (size_t)B.npairs, L.stream));
  L.timed = L.timed2 = false;
  if (p.use_packed) {
    TB_CUDA(ctx, cudaEventRecord(L.k0, L.stream));
    B.counter = counters;           // [0] queue head, [1] pairs completed by the packed kernel
    TB_CUDA(ctx, tb::launch_gotoh_packed(p.tbmode, 4, B, p.blocks_packed, L.stream));
    TB_CUDA(ctx, cudaEventRecord(L.k1, L.stream));
    L.timed = true;
    L.kend = L.k1;
  } else if (p.use_pp) {
    TB_CUDA(ctx, cudaEventRecord(L.k0, L.stream));
    B.counter = counters;           // [0] ticket counter, [1] pairs completed by the 4-channel kernel
    TB_CUDA(ctx, tb::launch_gotoh_pp(4, traceback, p.pp_harr, B, L.ppw, p.blocks_pp, L.stream));
    TB_CUDA(ctx, cudaEventRecord(L.k1, L.stream));
    L.timed = true;
    L.kend = L.k1;
  } else {
    TB_CUDA(ctx, cudaEventRecord(L.k1, L.stream));
    B.counter = counters + 8;
    TB_CUDA(ctx, tb::launch_gotoh_general(mode, traceback, B, p.blocks_general, L.stream));
    TB_CUDA(ctx, cudaEventRecord(L.k2, L.stream));
    L.timed2 = true;
    L.kend = L.k2;
  }
  ctx->launches++;
  TB_CUDA(ctx, L.cnt.reserve(64));
  TB_CUDA(ctx, cudaMemcpyAsync(L.cnt.p, counters, 64, cudaMemcpyDeviceToHost, L.stream));
  L.view = B; L.plan = p; L.mode = mode; L.traceback = traceback;
  return TB_OK;
}

int collect_timing(tb_ctx* ctx, Lane& L);

// After the lane's stream has drained: run stage 2 when the packed kernel left pairs behind. *ran says whether it did
// (the caller then repeats its D2H of the results).
int finish_gotoh(tb_ctx* ctx, Lane& L, bool* ran) {
  *ran = false;
  if ((!L.plan.use_packed && !L.plan.use_pp) || !L.timed) return TB_OK;
  const unsigned int* cnt = static_cast<const unsigned int*>(L.cnt.p);
  if (cnt[1] >= (unsigned)L.view.npairs) return TB_OK;
  if (int rc = collect_timing(ctx, L)) return rc;          // stage 1's events are about to be re-recorded
  unsigned int* counters = L.counter.as<unsigned int>();
  tb::GotohBatch B = L.view;
  TB_CUDA(ctx, cudaEventRecord(L.k0, L.stream));
  B.counter = counters + 2;       // second queue head; its completion count lands in counters[3]
  if (L.plan.use_pp) TB_CUDA(ctx, tb::launch_gotoh_pp(5, L.traceback, L.plan.pp_harr, B, L.ppw, L.plan.blocks_pp5, L.stream));
  else TB_CUDA(ctx, tb::launch_gotoh_packed(L.plan.tbmode, 5, B, L.plan.blocks_packed5, L.stream));
  TB_CUDA(ctx, cudaEventRecord(L.k1, L.stream));
  B.counter = counters + 8;
  TB_CUDA(ctx, tb::launch_gotoh_general(L.mode, L.traceback, B, L.plan.blocks_general, L.stream));
  TB_CUDA(ctx, cudaEventRecord(L.k2, L.stream));
  ctx->launches += 2;
  L.timed = L.timed2 = true;
  L.kend = L.k2;
  TB_CUDA(ctx, cudaMemsetAsync(counters + 1, 0, 4, L.stream));   // stage 1's count is already in last_packed_pairs
  TB_CUDA(ctx, cudaMemcpyAsync(L.cnt.p, counters, 64, cudaMemcpyDeviceToHost, L.stream));
  *ran = true;
  return TB_OK;
}

int collect_timing(tb_ctx* ctx, Lane& L) {
  float ms = 0;
  if (L.timed) { TB_CUDA(ctx, cudaEventElapsedTime(&ms, L.k0, L.k1)); ctx->last_fast_ms += ms; }
  if (L.timed2) { TB_CUDA(ctx, cudaEventElapsedTime(&ms, L.k1, L.k2)); ctx->last_general_ms += ms; }
  if (L.timed && L.cnt.p) ctx->last_packed_pairs += static_cast<const unsigned int*>(L.cnt.p)[1] + static_cast<const unsigned int*>(L.cnt.p)[3];
  L.timed = L.timed2 = false;
  return TB_OK;
}

size_t elem_size_a(int mode) { return mode == tb::kModeSS ? 1 : sizeof(float); }
size_t elem_size_b(int mode) { return mode == tb::kModePP ? sizeof(float) : 1; }
long long item_elems_a(int mode, int len) { return mode == tb::kModeSS ? len : 6ll * len; }
long long item_elems_b(int mode, int len) { return mode == tb::kModePP ? 6ll * len : len; }

int check_range(tb_ctx* ctx, const Shape& sh, tb_score sc) {
  const long long maxabs = std::max(std::max(llabs((long long)sc.match), llabs((long long)sc.mismatch)),
                                    std::max(llabs((long long)sc.gap_open), llabs((long long)sc.gap_extend)));
  const long long bound = 2 * llabs((long long)sc.gap_open) + ((long long)sh.maxsum + 2) * maxabs;
  if (bound >= 900000)
    return fail(ctx, TB_ERR_UNSUPPORTED, "score range reaches the reference's inf sentinel (1e6): sizes x scores too large");
  return TB_OK;
}

int run_gotoh(tb_ctx* ctx, int mode, const tb_batch* batch, tb_score sc, tb_align_config ac, tb_result* res) {
  if (!ctx) return TB_ERR_INVALID;
  if (!batch || !res) return fail(ctx, TB_ERR_INVALID, "null batch/result");
  const size_t np = batch->npairs;
  ctx->last_fast_ms = ctx->last_general_ms = ctx->last_call_ms = 0;
  ctx->last_packed_pairs = 0;
  if (np == 0) return TB_OK;
  if (np > (size_t)INT_MAX) return fail(ctx, TB_ERR_INVALID, "npairs too large");
  if (!batch->a1.base || !batch->a1.off || !batch->a1.len || !batch->a2.base || !batch->a2.off || !batch->a2.len || !res->scores)
    return fail(ctx, TB_ERR_INVALID, "null arena/score pointer");
  const bool want_rows = res->row0 != nullptr || res->row1 != nullptr;
  if (want_rows && (!res->row0 || !res->row1)) return fail(ctx, TB_ERR_INVALID, "row0 and row1 go together");
  const bool traceback = res->ops != nullptr || want_rows;
  const bool packed_ops = res->ops != nullptr && res->ops_packed != 0;
  if (traceback && !res->ops_len) return fail(ctx, TB_ERR_INVALID, "ops / rows given without ops_len");
  TB_CUDA(ctx, cudaSetDevice(ctx->device));

  // Lengths on the host (needed for validation and scratch sizing in both memory modes).
  const int32_t *l1 = batch->a1.len, *l2 = batch->a2.len;
  const int32_t mem = batch->mem & 0xff;
  const bool a1_trace_profiles = (batch->mem & TB_A1_TRACE_PROFILES) != 0;   // the caller vouches: rows 4 (N) and 5 ('-') of every a1 profile are zero
  if (mem == TB_MEM_DEVICE) {
    ctx->tmp_len1.resize(np); ctx->tmp_len2.resize(np);
    TB_CUDA(ctx, cudaMemcpy(ctx->tmp_len1.data(), batch->a1.len, np * 4, cudaMemcpyDeviceToHost));
    TB_CUDA(ctx, cudaMemcpy(ctx->tmp_len2.data(), batch->a2.len, np * 4, cudaMemcpyDeviceToHost));
    l1 = ctx->tmp_len1.data(); l2 = ctx->tmp_len2.data();
  } else if (mem != TB_MEM_HOST) {
    return fail(ctx, TB_ERR_INVALID, "batch->mem must be TB_MEM_HOST or TB_MEM_DEVICE");
  }
  for (size_t i = 0; i < np; ++i)
    if (l1[i] < 0 || l2[i] < 0) return fail(ctx, TB_ERR_INVALID, "negative length");
  // profile x profile: which pairs are spread over many warps (per-pair scratch instead of per-slot scratch)
  std::vector<uint8_t> big;
  size_t pp_tickets = 0;
  ctx->last_big_pairs = 0;
  if (mode == tb::kModePP && getenv("TRACY_B200_NO_PPFAST") == nullptr && getenv("TRACY_B200_NO_BIG") == nullptr) {
    if (ctx->free_at_first_plan == 0) {
      size_t fr = 0, tot = 0;
      TB_CUDA(ctx, cudaMemGetInfo(&fr, &tot));
      ctx->free_at_first_plan = fr;
    }
    const size_t budget = (ctx->scratch_limit ? ctx->scratch_limit : ctx->free_at_first_plan / 3) / 2;
    pp_tickets = classify_pp(ctx, l1, l2, np, traceback, budget, big);
  }
  const uint8_t* bigp = big.empty() ? nullptr : big.data();
  Shape all;
  accumulate(all, l1, l2, np, mode != tb::kModePP, bigp);
  if (res->ops && !packed_ops && res->ops_stride < all.maxsum) return fail(ctx, TB_ERR_INVALID, "ops_stride smaller than max(len1+len2)");
  if (packed_ops && res->ops_stride < (all.maxsum + 3) / 4) return fail(ctx, TB_ERR_INVALID, "ops_stride smaller than ceil(max(len1+len2) / 4) bytes of packed ops");
  if (want_rows && res->rows_stride < all.maxsum) return fail(ctx, TB_ERR_INVALID, "rows_stride smaller than max(len1+len2)");
  if (int rc = check_range(ctx, all, sc)) return rc;
  // the DP kernels write one byte per op: into the caller's buffer when that is the form asked for, else into a lane buffer
  const bool plain_ops = res->ops != nullptr && !packed_ops;
  const int64_t ustride = plain_ops ? res->ops_stride : (((int64_t)all.maxsum + 15) & ~(int64_t)15);

  tb::GotohBatch B{};
  B.match = sc.match; B.mismatch = sc.mismatch; B.go = sc.gap_open; B.ge = sc.gap_extend;
  B.hfree = ac.h_free != 0; B.vfree = ac.v_free != 0;
  B.a_is_seq = mode == tb::kModeSS;
  B.mode = mode;
  B.order = nullptr;

  if (mem == TB_MEM_DEVICE) {
    Lane& L = ctx->lanes[0];
    Plan plan;
    if (int rc = make_plan(ctx, mode, traceback, all, np, sc, &plan, pp_tickets)) return rc;
    if (int rc = reserve_scratch(ctx, L, plan)) return rc;
    if (plan.use_pp) if (int rc = build_pp_work(ctx, L, l1, l2, bigp, np, traceback)) return rc;
    TB_CUDA(ctx, L.status.reserve(np));
    B.a_base = batch->a1.base; B.a_off = batch->a1.off; B.a_len = batch->a1.len;
    B.b_base = batch->a2.base; B.b_off = batch->a2.off; B.b_len = batch->a2.len;
    if (traceback && !plain_ops) TB_CUDA(ctx, L.ops.reserve(np * (size_t)ustride));
    B.scores = res->scores; B.ops = !traceback ? nullptr : plain_ops ? res->ops : L.ops.as<uint8_t>(); B.ops_stride = ustride; B.ops_len = res->ops_len;
    B.status = L.status.as<uint8_t>(); B.npairs = (int)np;
    B.row0 = res->row0; B.row1 = res->row1; B.rows_stride = res->rows_stride;
    B.opk = packed_ops ? res->ops : nullptr; B.opk_stride = packed_ops ? res->ops_stride : 0;
    if (int rc = enqueue_gotoh(ctx, L, mode, traceback, B, plan)) return rc;
    TB_CUDA(ctx, cudaStreamSynchronize(L.stream));
    bool again = false;
    if (int rc = finish_gotoh(ctx, L, &again)) return rc;
    if (again) TB_CUDA(ctx, cudaStreamSynchronize(L.stream));
    TB_CUDA(ctx, cudaEventElapsedTime(&ctx->last_call_ms, L.c0, L.kend));
    return collect_timing(ctx, L);
  }

  // ---- TB_MEM_HOST: chunks pipelined over kLanes streams, each H2D -> kernels -> D2H ----
  Plan plan;
  if (int rc = make_plan(ctx, mode, traceback, all, np, sc, &plan, pp_tickets)) return rc;
  // A chunk is a whole number of waves of the persistent grid (pairs of one batch cost about the same, so a
  // chunk then ends without a tail of idle warp slots), about an eighth of the batch, inputs <= 768 MiB.
  const size_t wave = std::max<size_t>(plan.slots, 1);
  // steady chunks of about a sixteenth of the batch (3 waves at 100 k pairs): measured best of 3 .. 21 waves once the ramps are in
  // (profiles/r02_e2e_chunk_probe.json: 89.3 ms against 91.4 - 91.9 ms; the chunk size hardly matters -- the next kernel's blocks
  // move into the slots the previous kernel's blocks leave, what a boundary costs is the warps of a block idling until its
  // slowest one is through, which the streamed form below pays only once -- and small chunks keep the copies short)
  size_t chunk_target = 7104, chunk_parts = 16;
  if (const char* ce = getenv("TRACY_B200_CHUNK_TARGET")) { const long v = atol(ce); if (v > 0) chunk_target = (size_t)v; }
  if (const char* ce = getenv("TRACY_B200_CHUNK_PARTS")) { const long v = atol(ce); if (v > 0) chunk_parts = (size_t)v; }
  size_t chunk = std::max<size_t>(2 * wave, std::min<size_t>(chunk_target, (np + chunk_parts - 1) / chunk_parts));
  chunk = std::max<size_t>(1, chunk / wave) * wave;
  {
    const double per_pair = (double)item_elems_a(mode, all.maxm) * elem_size_a(mode) + (double)item_elems_b(mode, all.maxn) * elem_size_b(mode) +
                            (traceback ? (double)ustride : 0.0);
    const size_t cap = (size_t)std::max(1.0, (768.0 * 1024 * 1024) / std::max(per_pair, 1.0));
    if (chunk > cap) chunk = cap >= wave ? cap / wave * wave : cap;
  }
  chunk = std::min(chunk, np);
  bool ramp = chunk >= 4 * wave;   // the first chunks are 1, 2 and 4 waves so that the kernels start after a short first copy
  {
    // Pooled arenas (an all-pairs list is two index arrays into ONE set of profiles): when everything the call touches and
    // returns is small, there is nothing to pipeline -- one chunk, one launch, instead of re-sending the pool with every chunk.
    long long amin = LLONG_MAX, amax = 0, bmin = LLONG_MAX, bmax = 0;
    for (size_t i = 0; i < np; ++i) {
      const long long ao = batch->a1.off[i], bo = batch->a2.off[i];
      amin = std::min(amin, ao); amax = std::max(amax, ao + item_elems_a(mode, l1[i]));
      bmin = std::min(bmin, bo); bmax = std::max(bmax, bo + item_elems_b(mode, l2[i]));
    }
    const double touched = (double)(amax - amin) * elem_size_a(mode) + (double)(bmax - bmin) * elem_size_b(mode) +
                           (double)np * (28.0 + (traceback ? (double)ustride : 0.0) + (want_rows ? 2.0 * (double)res->rows_stride : 0.0));
    if (amin >= 0 && bmin >= 0 && touched <= 512.0 * 1024 * 1024) { chunk = np; ramp = false; }
  }
  if (const char* ce = getenv("TRACY_B200_CHUNK")) { const long v = atol(ce); if (v > 0) { chunk = std::min<size_t>((size_t)v, np); ramp = false; } }   // tuning knob
  const bool trace = getenv("TRACY_B200_TRACE") != nullptr;   // per-chunk timeline on stderr (profiles/)
  const size_t esa = elem_size_a(mode), esb = elem_size_b(mode);
  // Profile rows 5 ('-') never enter _score (k < 5, src/align.h:112-116): when the a1 items are equally long and
  // back to back, copy rows 0..4 of each item only (one strided DMA), a sixth less H2D traffic.
  bool a_rows5 = mode != tb::kModeSS && np > 1 && getenv("TRACY_B200_FULL_ROWS") == nullptr;
  for (size_t i = 1; a_rows5 && i < np; ++i)
    a_rows5 = l1[i] == l1[0] && batch->a1.off[i] == batch->a1.off[0] + (int64_t)i * 6 * l1[0];
  if (a_rows5 && l1[0] == 0) a_rows5 = false;
  const bool no_rows4 = getenv("TRACY_B200_NO_ROWS4") != nullptr;
  int nlanes = 3;
  if (const char* le = getenv("TRACY_B200_LANES")) nlanes = std::max(1, std::min(kLanes, atoi(le)));   // tuning knob
  TB_CUDA(ctx, cudaEventRecord(ctx->t0, ctx->lanes[0].stream));
  for (int i = 0; i < kLanes; ++i) ctx->lanes[i].chunk = -1;
  auto copy_out = [&](Lane& L, size_t p0, size_t cn) -> int {
    TB_CUDA(ctx, cudaMemcpyAsync(res->scores + p0, L.scores.p, cn * 4, cudaMemcpyDeviceToHost, L.stream));
    ctx->d2h += cn * 4;
    if (traceback) {
      if (plain_ops) TB_CUDA(ctx, cudaMemcpyAsync(res->ops + p0 * (size_t)ustride, L.ops.p, cn * (size_t)ustride, cudaMemcpyDeviceToHost, L.stream));
      if (packed_ops) TB_CUDA(ctx, cudaMemcpyAsync(res->ops + p0 * (size_t)res->ops_stride, L.opk.p, cn * (size_t)res->ops_stride, cudaMemcpyDeviceToHost, L.stream));
      if (want_rows) {
        TB_CUDA(ctx, cudaMemcpyAsync(res->row0 + p0 * (size_t)res->rows_stride, L.row0.p, cn * (size_t)res->rows_stride, cudaMemcpyDeviceToHost, L.stream));
        TB_CUDA(ctx, cudaMemcpyAsync(res->row1 + p0 * (size_t)res->rows_stride, L.row1.p, cn * (size_t)res->rows_stride, cudaMemcpyDeviceToHost, L.stream));
      }
      TB_CUDA(ctx, cudaMemcpyAsync(res->ops_len + p0, L.ops_len.p, cn * 4, cudaMemcpyDeviceToHost, L.stream));
      ctx->d2h += (plain_ops ? cn * (size_t)ustride : 0) + (packed_ops ? cn * (size_t)res->ops_stride : 0) + (want_rows ? 2 * cn * (size_t)res->rows_stride : 0) + cn * 4;
    }
    return TB_OK;
  };
  auto retire = [&](Lane& L) -> int {   // wait for the lane's chunk, fold its timing in
    TB_CUDA(ctx, cudaStreamSynchronize(L.stream));
    if (L.chunk >= 0) {
      bool again = false;
      if (int rc = finish_gotoh(ctx, L, &again)) return rc;
      if (again) {                      // pairs the packed kernel left: stage 2 ran, fetch the results again
        if (int rc = copy_out(L, L.p0, (size_t)L.view.npairs)) return rc;
        if (trace) TB_CUDA(ctx, cudaEventRecord(L.d1, L.stream));
        TB_CUDA(ctx, cudaStreamSynchronize(L.stream));
      }
    }
    if (trace && L.chunk >= 0) {
      float th0 = 0, tk0 = 0, tk2 = 0, td1 = 0;
      cudaEventElapsedTime(&th0, ctx->t0, L.h0); cudaEventElapsedTime(&tk0, ctx->t0, L.timed ? L.k0 : L.k1);
      cudaEventElapsedTime(&tk2, ctx->t0, L.kend); cudaEventElapsedTime(&td1, ctx->t0, L.d1);
      fprintf(stderr, "tracy_b200 chunk %ld lane %d: h2d %.2f..%.2f kernels ..%.2f d2h ..%.2f ms\n", L.chunk, (int)(&L - ctx->lanes), th0, tk0, tk2, td1);
    }
    L.chunk = -1;
    return collect_timing(ctx, L);
  };
  // The chunk pipeline as one unit: on ANY error the lanes still hold kernels and D2H copies into the caller's buffers, so they
  // are drained before the error is returned (the caller may free those buffers right away).
  // The chunk schedule. Ramp up: 1, 2, 4 waves, so the first kernel starts after a sub-millisecond copy. Ramp down: the last
  // chunks are 4, 2, 1 waves, so what is left after the last kernel is the device-to-host copy of ONE wave's results. The steady
  // chunks in between share what remains evenly in whole waves (no lone remainder chunk running on a mostly idle GPU).
  std::vector<size_t> sched;
  if (ramp && np >= 30 * wave && getenv("TRACY_B200_NO_RAMPDOWN") == nullptr) {
    const size_t head[3] = {wave, 2 * wave, 4 * wave}, tail_w[3] = {4 * wave, 2 * wave, wave};
    const size_t rest = np - 7 * wave;                    // pairs left for the steady chunks and the ramp down
    const size_t tail_sum = 7 * wave + (rest % wave);    // the ramp down also takes the part of a wave at the very end
    const size_t mid = rest - tail_sum;                   // a whole number of waves
    const size_t nmid = std::max<size_t>(1, (mid + chunk - 1) / chunk);
    for (size_t h : head) sched.push_back(h);
    for (size_t k = 0; k < nmid; ++k) {                   // spread the middle waves over nmid chunks
      const size_t waves = mid / wave, w = waves / nmid + (k < waves % nmid ? 1 : 0);
      if (w) sched.push_back(w * wave);
    }
    sched.push_back(tail_w[0]); sched.push_back(tail_w[1]); sched.push_back(tail_w[2] + rest % wave);
  } else {
    for (size_t ci = 0, p0 = 0; p0 < np; ++ci) { const size_t cn = std::min(ramp && ci < 3 ? (wave << ci) : chunk, np - p0); sched.push_back(cn); p0 += cn; }
  }
  // ---- Streamed form: ONE launch of the packed kernel for the whole batch (GotohBatch::gate_*, gotoh_packed.cu). The copy
  // stream brings the chunks in one after the other and raises the count of pairs that have landed after each; the kernel's
  // warps take pairs in index order and wait at that count; the warp that finishes the last pair of a chunk sets the chunk's
  // word in page-locked host memory, and the host then sends that chunk's results off on a third stream. What a launch per
  // chunk loses at every boundary -- the warps of a block idling until its slowest one is through -- is paid once, at the
  // end. Needs the whole batch resident (2.4 GB in, 1.1 GB out for 100 k pairs of 1000 x 4000: HBM has room), chunks that
  // occupy disjoint, increasing arena ranges, and the packed kernel; anything else takes the chunk pipeline below.
  const auto call_t0 = std::chrono::steady_clock::now();
  auto streamed = [&](bool* took) -> int {
    *took = false;
    auto decline = [&](const char* why) { if (trace) fprintf(stderr, "tracy_b200 streamed form not taken: %s\n", why); return TB_OK; };
    if (getenv("TRACY_B200_NO_STREAM")) return TB_OK;
    if (!plan.use_packed) return decline("not a packed-kernel batch");
    if (np < 16 * wave) return decline("fewer than 16 waves");
    // When the host link is what binds (eight processes pulling through one root complex), the launch-per-chunk pipeline
    // moves the same bytes ~6 % sooner (8 x B200: 172 against 183 ms per step): a streamed call whose last input landed
    // in the last 15 % of the kernel's run sends the next 31 eligible calls down the other road.
    const bool force_stream = getenv("TRACY_B200_FORCE_STREAM") != nullptr;   // tests: always streamed when eligible
    bool inputs_pinned = false;
    {
      cudaPointerAttributes pa{}, pb{};
      inputs_pinned = cudaPointerGetAttributes(&pa, batch->a1.base) == cudaSuccess && pa.type == cudaMemoryTypeHost &&
                      cudaPointerGetAttributes(&pb, batch->a2.base) == cudaSuccess && pb.type == cudaMemoryTypeHost;
      cudaGetLastError();
    }
    if (!inputs_pinned && (getenv("CUDA_LAUNCH_BLOCKING") || getenv("CUDA_INJECTION64_PATH") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR")))
      return decline("pageable inputs go in after the launch, and this process may not return from a launch before the kernel ends");
    if (ctx->stream_backoff > 0 && !force_stream) { --ctx->stream_backoff; return decline("the host link was the limit last time: launch per chunk"); }
    // its own chunk schedule, in whole waves: 1, 1, 2, 4 to start the kernel after a short first copy, 4 in the middle, 2 and 1 (plus
    // the part of a wave at the very end) so that little is left to send when the kernel ends
    std::vector<size_t> sched;
    {
      const size_t rem = np % wave, waves = np / wave;    // waves >= 16
      for (size_t w : {(size_t)1, (size_t)1, (size_t)2, (size_t)4}) sched.push_back(w * wave);
      size_t mid = waves - 8 - 3;
      while (mid >= 8) { sched.push_back(4 * wave); mid -= 4; }
      if (mid) sched.push_back(mid * wave);
      sched.push_back(2 * wave); sched.push_back(wave + rem);
    }
    const size_t nch = sched.size();
    std::vector<long long> ca0(nch), ca1(nch), cb0(nch), cb1(nch);
    std::vector<size_t> cp0(nch + 1, 0);
    for (size_t c = 0, p0 = 0; c < nch; ++c) {
      const size_t cn = std::min(sched[c], np - p0);
      long long amin = LLONG_MAX, amax = LLONG_MIN, bmin = LLONG_MAX, bmax = LLONG_MIN;
      for (size_t i = p0; i < p0 + cn; ++i) {
        const long long ao = batch->a1.off[i], bo = batch->a2.off[i];
        if (ao < 0 || bo < 0) return fail(ctx, TB_ERR_INVALID, "negative arena offset");
        amin = std::min(amin, ao); amax = std::max(amax, ao + item_elems_a(mode, l1[i]));
        bmin = std::min(bmin, bo); bmax = std::max(bmax, bo + item_elems_b(mode, l2[i]));
      }
      if (c && (amin < ca1[c - 1] || bmin < cb1[c - 1])) return decline("chunks share arena ranges");   // or use them out of order: chunk pipeline
      ca0[c] = amin; ca1[c] = amax; cb0[c] = bmin; cb1[c] = bmax;
      p0 += cn; cp0[c + 1] = p0;
    }
    if (cp0[nch] != np) return decline("schedule does not cover the batch");
    const long long a_lo = ca0[0], b_lo = cb0[0];
    const size_t abytes = (size_t)(ca1[nch - 1] - a_lo) * esa, bbytes = (size_t)(cb1[nch - 1] - b_lo) * esb;
    const size_t out_per_pair = 4 + 1 + (traceback ? (size_t)ustride + 4 : 0) + (want_rows ? 2 * (size_t)res->rows_stride : 0) + (packed_ops ? (size_t)res->ops_stride : 0);
    if (ctx->lanes[0].a.cap < abytes + 16 || ctx->lanes[0].b.cap < bbytes + 16 || ctx->lanes[0].row0.cap < (want_rows ? np * (size_t)res->rows_stride : 0) ||
        ctx->lanes[0].ops.cap < (traceback ? np * (size_t)ustride : 0)) {
      size_t fr = 0, tot = 0;
      TB_CUDA(ctx, cudaMemGetInfo(&fr, &tot));
      const size_t held = ctx->lanes[0].a.cap + ctx->lanes[0].b.cap + ctx->lanes[0].ops.cap + ctx->lanes[0].row0.cap + ctx->lanes[0].row1.cap + ctx->lanes[0].opk.cap;
      if (abytes + bbytes + np * (out_per_pair + 24) > (fr + held) / 2) return decline("batch does not fit in half of the free HBM");
    }
    Lane& K = ctx->lanes[0];                              // kernel stream; lanes[1].stream: copies in, lanes[2].stream: copies out
    cudaStream_t s_in = ctx->lanes[1].stream, s_out = ctx->lanes[2].stream;
    if (int rc = reserve_scratch(ctx, K, plan)) return rc;
    TB_CUDA(ctx, K.a.reserve(abytes + 16)); TB_CUDA(ctx, K.b.reserve(bbytes + 16));
    TB_CUDA(ctx, K.scores.reserve(np * 4)); TB_CUDA(ctx, K.status.reserve(np));
    if (traceback) { TB_CUDA(ctx, K.ops.reserve(np * (size_t)ustride)); TB_CUDA(ctx, K.ops_len.reserve(np * 4)); }
    if (want_rows) { TB_CUDA(ctx, K.row0.reserve(np * (size_t)res->rows_stride)); TB_CUDA(ctx, K.row1.reserve(np * (size_t)res->rows_stride)); }
    if (packed_ops) TB_CUDA(ctx, K.opk.reserve(np * (size_t)res->ops_stride));
    TB_CUDA(ctx, K.meta.reserve(np * 24)); TB_CUDA(ctx, K.meta_d.reserve(np * 24));
    // gate words: device [ready | pad .. 64 | done[nch] | end[nch]], host (page-locked) [flag[nch] | ready value per chunk [nch] | end[nch]]
    const size_t gate_dev = 64 + nch * 8;
    TB_CUDA(ctx, K.gate.reserve(gate_dev)); TB_CUDA(ctx, K.gate_host.reserve(nch * 12));
    volatile int32_t* hflag = static_cast<volatile int32_t*>(K.gate_host.p);
    uint32_t* hready = reinterpret_cast<uint32_t*>(const_cast<int32_t*>(hflag) + nch);
    int32_t* hend = reinterpret_cast<int32_t*>(hready + nch);
    for (size_t c = 0; c < nch; ++c) { hflag[c] = 0; hready[c] = (uint32_t)cp0[c + 1]; hend[c] = (int32_t)cp0[c + 1]; }
    void* hflag_dev = nullptr;
    TB_CUDA(ctx, cudaHostGetDevicePointer(&hflag_dev, K.gate_host.p, 0));
    unsigned int* g_ready = K.gate.as<unsigned int>();
    unsigned int* g_done = g_ready + 16;
    int32_t* g_end = reinterpret_cast<int32_t*>(g_done + nch);

    int64_t* hoff = static_cast<int64_t*>(K.meta.p);
    int32_t* hlen = reinterpret_cast<int32_t*>(hoff + 2 * np);
    for (size_t i = 0; i < np; ++i) { hoff[i] = batch->a1.off[i] - a_lo; hoff[np + i] = batch->a2.off[i] - b_lo; }
    std::memcpy(hlen, l1, np * 4); std::memcpy(hlen + np, l2, np * 4);

    TB_CUDA(ctx, cudaEventRecord(ctx->t0, K.stream));
    // Nothing that runs as a kernel may be queued behind the launch on the copy streams: the persistent grid holds every SM
    // while it waits at the gate, and a memset queued ahead of the gate's update would never start. Rows 4 and 5 of a
    // 4-row upload are therefore zeroed for the whole batch up front.
    const int up_rows = !a_rows5 ? 6 : (!no_rows4 && a1_trace_profiles) ? 4 : want_rows ? 6 : 5;
    if (a_rows5 && up_rows == 4) {
      const size_t len = (size_t)l1[0], pitch = 6 * len * 4;
      TB_CUDA(ctx, cudaMemset2DAsync((char*)K.a.p + 4 * len * 4, pitch, 0, 2 * len * 4, np, s_in));
    }
    TB_CUDA(ctx, cudaMemsetAsync(K.gate.p, 0, gate_dev, s_in));
    TB_CUDA(ctx, cudaMemcpyAsync(g_end, hend, nch * 4, cudaMemcpyHostToDevice, s_in));
    TB_CUDA(ctx, cudaMemcpyAsync(K.meta_d.p, hoff, np * 24, cudaMemcpyHostToDevice, s_in));
    TB_CUDA(ctx, cudaEventRecord(K.in_done, s_in));
    TB_CUDA(ctx, cudaStreamWaitEvent(K.stream, K.in_done, 0));
    ctx->h2d += np * 24 + nch * 4;

    // Inputs, chunk by chunk, each followed by the new count. Page-locked inputs: every copy is queued BEFORE the launch, so
    // the kernel's progress never depends on what this thread does after it -- a launch call that does not return until the
    // kernel has finished (a profiler serialising launches, CUDA_LAUNCH_BLOCKING) then costs the overlap, not the call.
    // Pageable inputs block in cudaMemcpyAsync: they are sent after the launch, results leaving in between.
    auto send_in = [&](size_t c) -> int {
      const size_t p0 = cp0[c], cn = cp0[c + 1] - p0;
      const size_t ab = (size_t)(ca1[c] - ca0[c]) * esa, bb = (size_t)(cb1[c] - cb0[c]) * esb;
      char* da = (char*)K.a.p + (size_t)(ca0[c] - a_lo) * esa;
      if (a_rows5) {
        const size_t len = (size_t)l1[0], pitch = 6 * len * 4;
        const float* src = (const float*)batch->a1.base + ca0[c];
        TB_CUDA(ctx, cudaMemcpy2DAsync(da, pitch, src, pitch, (size_t)up_rows * len * 4, cn, cudaMemcpyHostToDevice, s_in));
        ctx->h2d += (size_t)up_rows * len * 4 * cn;
      } else {
        TB_CUDA(ctx, cudaMemcpyAsync(da, (const char*)batch->a1.base + (size_t)ca0[c] * esa, ab, cudaMemcpyHostToDevice, s_in));
        ctx->h2d += ab;
      }
      TB_CUDA(ctx, cudaMemcpyAsync((char*)K.b.p + (size_t)(cb0[c] - b_lo) * esb, (const char*)batch->a2.base + (size_t)cb0[c] * esb, bb, cudaMemcpyHostToDevice, s_in));
      TB_CUDA(ctx, cudaMemcpyAsync(g_ready, hready + c, 4, cudaMemcpyHostToDevice, s_in));
      ctx->h2d += bb + 4;
      return TB_OK;
    };
    if (inputs_pinned) {
      for (size_t c = 0; c < nch; ++c) if (int rc = send_in(c)) return rc;
      TB_CUDA(ctx, cudaEventRecord(ctx->lanes[1].h0, s_in));
    }

    tb::GotohBatch C = B;
    int64_t* doff = K.meta_d.as<int64_t>();
    int32_t* dlen = reinterpret_cast<int32_t*>(doff + 2 * np);
    C.a_base = K.a.p; C.a_off = doff; C.a_len = dlen;
    C.b_base = K.b.p; C.b_off = doff + np; C.b_len = dlen + np;
    C.scores = K.scores.as<int32_t>(); C.ops = traceback ? K.ops.as<uint8_t>() : nullptr;
    C.ops_stride = ustride; C.ops_len = traceback ? K.ops_len.as<int32_t>() : nullptr;
    C.status = K.status.as<uint8_t>(); C.npairs = (int)np;
    C.row0 = want_rows ? K.row0.as<uint8_t>() : nullptr; C.row1 = want_rows ? K.row1.as<uint8_t>() : nullptr; C.rows_stride = res->rows_stride;
    C.opk = packed_ops ? K.opk.as<uint8_t>() : nullptr; C.opk_stride = packed_ops ? res->ops_stride : 0;
    C.gate_ready = g_ready; C.gate_done = g_done; C.gate_end = g_end; C.gate_host = static_cast<volatile int32_t*>(hflag_dev);
    if (int rc = enqueue_gotoh(ctx, K, mode, traceback, C, plan)) return rc;
    K.view.gate_ready = nullptr; K.view.gate_done = nullptr; K.view.gate_end = nullptr; K.view.gate_host = nullptr;   // a second stage runs ungated
    *took = true;

    auto send_out = [&](size_t p0, size_t cn) -> int {
      TB_CUDA(ctx, cudaMemcpyAsync(res->scores + p0, K.scores.as<int32_t>() + p0, cn * 4, cudaMemcpyDeviceToHost, s_out));
      ctx->d2h += cn * 4;
      if (traceback) {
        if (plain_ops) TB_CUDA(ctx, cudaMemcpyAsync(res->ops + p0 * (size_t)ustride, K.ops.as<uint8_t>() + p0 * (size_t)ustride, cn * (size_t)ustride, cudaMemcpyDeviceToHost, s_out));
        if (packed_ops) TB_CUDA(ctx, cudaMemcpyAsync(res->ops + p0 * (size_t)res->ops_stride, K.opk.as<uint8_t>() + p0 * (size_t)res->ops_stride, cn * (size_t)res->ops_stride, cudaMemcpyDeviceToHost, s_out));
        if (want_rows) {
          TB_CUDA(ctx, cudaMemcpyAsync(res->row0 + p0 * (size_t)res->rows_stride, K.row0.as<uint8_t>() + p0 * (size_t)res->rows_stride, cn * (size_t)res->rows_stride, cudaMemcpyDeviceToHost, s_out));
          TB_CUDA(ctx, cudaMemcpyAsync(res->row1 + p0 * (size_t)res->rows_stride, K.row1.as<uint8_t>() + p0 * (size_t)res->rows_stride, cn * (size_t)res->rows_stride, cudaMemcpyDeviceToHost, s_out));
        }
        TB_CUDA(ctx, cudaMemcpyAsync(res->ops_len + p0, K.ops_len.as<int32_t>() + p0, cn * 4, cudaMemcpyDeviceToHost, s_out));
        ctx->d2h += (plain_ops ? cn * (size_t)ustride : 0) + (packed_ops ? cn * (size_t)res->ops_stride : 0) + (want_rows ? 2 * cn * (size_t)res->rows_stride : 0) + cn * 4;
      }
      return TB_OK;
    };
    const auto w0 = std::chrono::steady_clock::now();
    auto wall_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count(); };
    if (trace) fprintf(stderr, "tracy_b200 streamed batch: %zu pairs in %zu chunks, launched %.2f ms into the call\n", np, nch,
                       std::chrono::duration<double, std::milli>(w0 - call_t0).count());
    auto after_launch = [&]() -> int {
    size_t sent = 0;                                      // chunks whose results are on their way
    auto drain = [&](bool wait) -> int {                  // send off every chunk the kernel has finished (wait: block for the next one)
      double idle0 = wall_ms();
      while (sent < nch) {
        if (!hflag[sent]) {
          if (!wait) return TB_OK;
          const cudaError_t q = cudaStreamQuery(K.stream);
          if (q != cudaSuccess && q != cudaErrorNotReady) { ctx->err = std::string("streamed batch: ") + cudaGetErrorString(q); cudaGetLastError(); return TB_ERR_CUDA; }
          if (q == cudaSuccess && !hflag[sent]) {         // the kernel is gone: its count is final (pairs left to a second stage are counted too)
            hflag[sent] = 1; continue;
          }
          if (wall_ms() - idle0 > 60000.0) return fail(ctx, TB_ERR_CUDA, "streamed batch: no chunk finished for 60 s");
          std::this_thread::yield();
          continue;
        }
        if (trace) fprintf(stderr, "tracy_b200 streamed chunk %zu (%zu pairs) finished at %.2f ms\n", sent, cp0[sent + 1] - cp0[sent], wall_ms());
        if (int rc = send_out(cp0[sent], cp0[sent + 1] - cp0[sent])) return rc;
        ++sent; idle0 = wall_ms();
      }
      return TB_OK;
    };
    if (!inputs_pinned) {
      for (size_t c = 0; c < nch; ++c) {
        if (int rc = send_in(c)) return rc;
        if (c + 1 == nch) TB_CUDA(ctx, cudaEventRecord(ctx->lanes[1].h0, s_in));
        if (int rc = drain(false)) return rc;
      }
    }
    if (int rc = drain(true)) return rc;
    TB_CUDA(ctx, cudaStreamSynchronize(K.stream));
    TB_CUDA(ctx, cudaStreamSynchronize(s_in));
    bool again = false;
    if (int rc = finish_gotoh(ctx, K, &again)) return rc;
    if (again) {                                          // pairs the packed kernel left: the second stage ran on the resident batch, fetch everything again
      TB_CUDA(ctx, cudaStreamSynchronize(K.stream));
      if (int rc = send_out(0, np)) return rc;
    }
    TB_CUDA(ctx, cudaStreamSynchronize(s_out));
    {
      float t_in = 0, t_k = 0;
      if (cudaEventElapsedTime(&t_in, ctx->t0, ctx->lanes[1].h0) == cudaSuccess && cudaEventElapsedTime(&t_k, ctx->t0, K.kend) == cudaSuccess &&
          t_in > 0.85f * t_k && !force_stream) ctx->stream_backoff = 31;
      cudaGetLastError();
      if (trace) fprintf(stderr, "tracy_b200 streamed batch: last input landed at %.2f ms, kernel done at %.2f ms, last result at %.2f ms%s\n", t_in, t_k, wall_ms(),
                         ctx->stream_backoff ? " -- link-bound: next calls launch per chunk" : "");
    }
    TB_CUDA(ctx, cudaEventElapsedTime(&ctx->last_call_ms, K.c0, K.kend));
    return collect_timing(ctx, K);
    };
    const int arc = after_launch();
    if (arc != TB_OK) {                                   // whatever went wrong, the kernel must not wait at the gate for ever
      const std::string keep = ctx->err;
      const uint32_t all = (uint32_t)np;
      cudaMemcpy(g_ready, &all, 4, cudaMemcpyHostToDevice);
      cudaGetLastError();
      ctx->err = keep;
    }
    return arc;
  };
  {
    bool took = false;
    const int src = streamed(&took);
    if (src != TB_OK || took) {
      if (src != TB_OK) {
        const std::string keep = ctx->err;
        if (ctx->lanes[0].gate.p) {                       // whatever failed: a kernel that may be waiting at the gate must be able to leave
          const uint32_t open = 0xffffffffu;
          cudaMemcpy(ctx->lanes[0].gate.p, &open, 4, cudaMemcpyHostToDevice);
        }
        for (int i = 0; i < kLanes; ++i) { cudaStreamSynchronize(ctx->lanes[i].stream); ctx->lanes[i].chunk = -1; ctx->lanes[i].timed = ctx->lanes[i].timed2 = false; }
        cudaGetLastError();
        ctx->err = keep;
      }
      return src;
    }
  }
  auto pipeline = [&]() -> int {
  size_t nchunks = 0;
  cudaEvent_t prev_in = nullptr;                          // inputs of the previous chunk have arrived
  for (size_t ci = 0, p0 = 0; p0 < np; ++ci, ++nchunks) {
    Lane& L = ctx->lanes[ci % nlanes];
    if (int rc = retire(L)) return rc;                    // chunk ci-nlanes is done; its buffers are free
    const size_t cn = std::min(sched[std::min(ci, sched.size() - 1)], np - p0);
    // extents of this chunk inside the caller's arenas
    long long amin = LLONG_MAX, amax = LLONG_MIN, bmin = LLONG_MAX, bmax = LLONG_MIN;
    for (size_t i = p0; i < p0 + cn; ++i) {
      const long long ao = batch->a1.off[i], bo = batch->a2.off[i];
      if (ao < 0 || bo < 0) return fail(ctx, TB_ERR_INVALID, "negative arena offset");
      amin = std::min(amin, ao); amax = std::max(amax, ao + item_elems_a(mode, l1[i]));
      bmin = std::min(bmin, bo); bmax = std::max(bmax, bo + item_elems_b(mode, l2[i]));
    }
    const size_t abytes = (size_t)(amax - amin) * esa, bbytes = (size_t)(bmax - bmin) * esb;
    const auto hc0 = std::chrono::steady_clock::now();
    auto hms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - hc0).count(); };
    Plan cp = plan;                                       // scratch per slot sized for the batch maxima, once
    if (cn < (size_t)plan.slots) {
      size_t tickets = cn;
      if (bigp) for (size_t i = p0; i < p0 + cn; ++i) if (bigp[i]) tickets += (size_t)pp_bands(l1[i]) - 1;
      if (int rc = make_plan(ctx, mode, traceback, all, cn, sc, &cp, tickets)) return rc;
    }
    const double h_plan = hms();
    if (int rc = reserve_scratch(ctx, L, cp)) return rc;
    const double h_scratch = hms();
    if (cp.use_pp) if (int rc = build_pp_work(ctx, L, l1 + p0, l2 + p0, bigp ? bigp + p0 : nullptr, cn, traceback)) return rc;
    const double h_pp = hms();
    TB_CUDA(ctx, L.a.reserve(abytes + 16)); TB_CUDA(ctx, L.b.reserve(bbytes + 16));
    TB_CUDA(ctx, L.scores.reserve(cn * 4)); TB_CUDA(ctx, L.status.reserve(cn));
    if (traceback) { TB_CUDA(ctx, L.ops.reserve(cn * (size_t)ustride)); TB_CUDA(ctx, L.ops_len.reserve(cn * 4)); }
    if (want_rows) { TB_CUDA(ctx, L.row0.reserve(cn * (size_t)res->rows_stride)); TB_CUDA(ctx, L.row1.reserve(cn * (size_t)res->rows_stride)); }
    if (packed_ops) TB_CUDA(ctx, L.opk.reserve(cn * (size_t)res->ops_stride));
    // offsets and lengths of the chunk go through ONE pinned block and one copy: [a_off | b_off | a_len | b_len]
    TB_CUDA(ctx, L.meta.reserve(cn * 24)); TB_CUDA(ctx, L.meta_d.reserve(cn * 24));
    if (trace && hms() > 2.0)
      fprintf(stderr, "tracy_b200 chunk %zu host: plan %.2f scratch %.2f pp work %.2f buffers %.2f ms (ptr %zu MB, pp_ptr %zu MB)\n", ci, h_plan, h_scratch - h_plan,
              h_pp - h_scratch, hms() - h_pp, L.ptr.cap >> 20, L.pp_ptr.cap >> 20);
    int64_t* hoff = static_cast<int64_t*>(L.meta.p);
    int32_t* hlen = reinterpret_cast<int32_t*>(hoff + 2 * cn);
    for (size_t i = 0; i < cn; ++i) { hoff[i] = batch->a1.off[p0 + i] - amin; hoff[cn + i] = batch->a2.off[p0 + i] - bmin; }
    std::memcpy(hlen, l1 + p0, cn * 4); std::memcpy(hlen + cn, l2 + p0, cn * 4);

    // one chunk's inputs at a time on the PCIe link: copies issued together on different streams share the bandwidth, which made the
    // FIRST kernel wait for a third of three copies (5 ms) instead of for its own 35 MB (0.7 ms)
    if (prev_in) TB_CUDA(ctx, cudaStreamWaitEvent(L.stream, prev_in, 0));
    if (trace) TB_CUDA(ctx, cudaEventRecord(L.h0, L.stream));
    if (a_rows5) {
      // rows 0..4 enter _score; row 5 ('-') only decides consensus characters (wanted with the gapped rows). Trace profiles made
      // by createProfile carry exact zeros in rows 4 and 5 (src/profile.h:37): when the caller says so (TB_A1_TRACE_PROFILES; a
      // host-side scan of 8 KB per pair would cost more than the copy it saves), 4 rows travel and the device copy's rows 4, 5
      // are zeroed in place -- 16 KB instead of 24 KB per 1 000-column profile.
      const size_t len = (size_t)l1[0], pitch = 6 * len * 4;
      const float* src = (const float*)batch->a1.base + amin;
      int rows = want_rows ? 6 : 5;
      if (!no_rows4 && a1_trace_profiles) rows = 4;
      TB_CUDA(ctx, cudaMemcpy2DAsync(L.a.p, pitch, src, pitch, (size_t)rows * len * 4, cn, cudaMemcpyHostToDevice, L.stream));
      if (rows == 4) TB_CUDA(ctx, cudaMemset2DAsync((char*)L.a.p + 4 * len * 4, pitch, 0, 2 * len * 4, cn, L.stream));
      ctx->h2d += (size_t)rows * len * 4 * cn;
    } else {
      TB_CUDA(ctx, cudaMemcpyAsync(L.a.p, (const char*)batch->a1.base + (size_t)amin * esa, abytes, cudaMemcpyHostToDevice, L.stream));
      ctx->h2d += abytes;
    }
    TB_CUDA(ctx, cudaMemcpyAsync(L.b.p, (const char*)batch->a2.base + (size_t)bmin * esb, bbytes, cudaMemcpyHostToDevice, L.stream));
    TB_CUDA(ctx, cudaMemcpyAsync(L.meta_d.p, hoff, cn * 24, cudaMemcpyHostToDevice, L.stream));
    TB_CUDA(ctx, cudaEventRecord(L.in_done, L.stream));
    prev_in = L.in_done;
    ctx->h2d += bbytes + cn * 24;

    tb::GotohBatch C = B;
    int64_t* doff = L.meta_d.as<int64_t>();
    int32_t* dlen = reinterpret_cast<int32_t*>(doff + 2 * cn);
    C.a_base = L.a.p; C.a_off = doff; C.a_len = dlen;
    C.b_base = L.b.p; C.b_off = doff + cn; C.b_len = dlen + cn;
    C.scores = L.scores.as<int32_t>(); C.ops = traceback ? L.ops.as<uint8_t>() : nullptr;
    C.ops_stride = ustride; C.ops_len = traceback ? L.ops_len.as<int32_t>() : nullptr;
    C.status = L.status.as<uint8_t>(); C.npairs = (int)cn;
    C.row0 = want_rows ? L.row0.as<uint8_t>() : nullptr; C.row1 = want_rows ? L.row1.as<uint8_t>() : nullptr; C.rows_stride = res->rows_stride;
    C.opk = packed_ops ? L.opk.as<uint8_t>() : nullptr; C.opk_stride = packed_ops ? res->ops_stride : 0;
    if (int rc = enqueue_gotoh(ctx, L, mode, traceback, C, cp)) return rc;
    L.chunk = (long)ci; L.p0 = p0;

    if (int rc = copy_out(L, p0, cn)) return rc;
    if (trace) TB_CUDA(ctx, cudaEventRecord(L.d1, L.stream));
    p0 += cn;
  }
  for (size_t k = 0; k < (size_t)nlanes; ++k)            // retire in issue order
    if (int rc = retire(ctx->lanes[(nchunks + k) % nlanes])) return rc;
  return TB_OK;
  };
  const int prc = pipeline();
  if (prc != TB_OK) {
    const std::string keep = ctx->err;
    for (int i = 0; i < kLanes; ++i) { cudaStreamSynchronize(ctx->lanes[i].stream); ctx->lanes[i].chunk = -1; ctx->lanes[i].timed = ctx->lanes[i].timed2 = false; }
    cudaGetLastError();
    ctx->err = keep;
  }
  return prc;
}

char cons_char(const float* p, int len, int pos) {
  // reference src/align.h:254-270: first strict maximum over the six rows, compared in double; >= 4 prints 'N'
  int best = 0;
  double bv = p[pos];
  for (int k = 1; k < 6; ++k) {
    const double v = p[(size_t)k * len + pos];
    if (v > bv) { bv = v; best = k; }
  }
  return best < 4 ? "ACGT"[best] : 'N';
}

}  // namespace

extern "C" {

const char* tb_version(void) { return "tracy_b200 0.1 (hot path of tracy v0.9.1; sm_100a)"; }

const char* tb_strerror(int code) {
  switch (code) {
    case TB_OK: return "ok";
    case TB_ERR_INVALID: return "invalid argument";
    case TB_ERR_CUDA: return "CUDA error / no usable device";
    case TB_ERR_NOMEM: return "out of memory";
    case TB_ERR_UNSUPPORTED: return "unsupported sizes or score values";
  }
  return "unknown error";
}
const char* tb_last_error(const tb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int tb_ctx_create(tb_ctx** out, int device) {
  if (!out) return TB_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) { cudaGetLastError(); return TB_ERR_CUDA; }
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return TB_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return TB_ERR_CUDA; }
  if (prop.major != 10) return TB_ERR_CUDA;   // sm_100a SASS only; there is no other code path
  tb_ctx* c = new (std::nothrow) tb_ctx();
  if (!c) return TB_ERR_NOMEM;
  c->device = device;
  c->sms = prop.multiProcessorCount;
  if (cudaEventCreate(&c->t0) != cudaSuccess) { cudaGetLastError(); delete c; return TB_ERR_CUDA; }
  {
    // The staging buffers of the small calls (profiles, basecalls, sweeps, fractions) come from the stream-ordered allocator; with the
    // default release threshold (0) the pool hands everything back at every synchronisation and each call pays for fresh device
    // memory again (a quarter of a second per basecall call of 256 traces). Keep up to 8 GB cached.
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
      unsigned long long keep = 8ull << 30;
      if (cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess) cudaGetLastError();
    } else cudaGetLastError();
  }
  for (int i = 0; i < kLanes; ++i) {
    Lane& L = c->lanes[i];
    if (cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&L.c0) != cudaSuccess ||
        cudaEventCreate(&L.k0) != cudaSuccess || cudaEventCreate(&L.h0) != cudaSuccess || cudaEventCreate(&L.d1) != cudaSuccess ||
        cudaEventCreate(&L.k1) != cudaSuccess || cudaEventCreate(&L.k2) != cudaSuccess ||
        cudaEventCreateWithFlags(&L.in_done, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      tb_ctx_destroy(c);
      return TB_ERR_CUDA;
    }
  }
  *out = c;
  return TB_OK;
}

void tb_ctx_destroy(tb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->t0) cudaEventDestroy(c->t0);
  for (int i = 0; i < kLanes; ++i) {
    Lane& L = c->lanes[i];
    if (L.stream) cudaStreamSynchronize(L.stream);
    DevBuf* bufs[] = {&L.a, &L.b, &L.meta_d, &L.scores, &L.ops, &L.ops_len, &L.status, &L.counter, &L.ptr, &L.rowbuf, &L.opsrev,
                      &L.pp_units, &L.pp_small, &L.pp_big, &L.pp_rowbuf, &L.pp_ptr, &L.pp_flags, &L.row0, &L.row1, &L.opk};
    for (DevBuf* b : bufs) b->release();
    L.meta.release(); L.cnt.release(); L.pp_stage.release();
    if (L.c0) cudaEventDestroy(L.c0);
    if (L.k0) cudaEventDestroy(L.k0);
    if (L.k1) cudaEventDestroy(L.k1);
    if (L.k2) cudaEventDestroy(L.k2);
    if (L.h0) cudaEventDestroy(L.h0);
    if (L.d1) cudaEventDestroy(L.d1);
    if (L.in_done) cudaEventDestroy(L.in_done);
    if (L.stream) cudaStreamDestroy(L.stream);
  }
  delete c;
}

int tb_host_alloc(tb_ctx* ctx, void** out, size_t bytes) {
  if (!ctx || !out) return TB_ERR_INVALID;
  TB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, TB_ERR_NOMEM, "cudaHostAlloc failed"); }
  return TB_OK;
}
int tb_host_free(tb_ctx* ctx, void* p) {
  if (!ctx) return TB_ERR_INVALID;
  if (p) TB_CUDA(ctx, cudaFreeHost(p));
  return TB_OK;
}
int tb_ctx_set_scratch_limit(tb_ctx* ctx, size_t bytes) {
  if (!ctx) return TB_ERR_INVALID;
  ctx->scratch_limit = bytes;
  return TB_OK;
}
int tb_ctx_stats(const tb_ctx* ctx, uint64_t* k, uint64_t* h2d, uint64_t* d2h) {
  if (!ctx) return TB_ERR_INVALID;
  if (k) *k = ctx->launches;
  if (h2d) *h2d = ctx->h2d;
  if (d2h) *d2h = ctx->d2h;
  return TB_OK;
}
int tb_ctx_last_kernel_ms(const tb_ctx* ctx, float* fast_ms, float* general_ms, float* sweep_ms) {
  if (!ctx) return TB_ERR_INVALID;
  if (fast_ms) *fast_ms = ctx->last_fast_ms;
  if (general_ms) *general_ms = ctx->last_general_ms;
  if (sweep_ms) *sweep_ms = ctx->last_sweep_ms;
  return TB_OK;
}

int tb_ctx_last_call_ms(const tb_ctx* ctx, float* device_ms) {
  if (!ctx || !device_ms) return TB_ERR_INVALID;
  *device_ms = ctx->last_call_ms;
  return TB_OK;
}

int tb_ctx_last_packed_pairs(const tb_ctx* ctx, uint64_t* pairs) {
  if (!ctx || !pairs) return TB_ERR_INVALID;
  *pairs = ctx->last_packed_pairs;
  return TB_OK;
}

int tb_ctx_last_big_pairs(const tb_ctx* ctx, uint64_t* pairs) {
  if (!ctx || !pairs) return TB_ERR_INVALID;
  *pairs = ctx->last_big_pairs;
  return TB_OK;
}

int tb_gotoh_ps(tb_ctx* ctx, const tb_batch* b, tb_score sc, tb_align_config ac, tb_result* r) { return run_gotoh(ctx, tb::kModePS, b, sc, ac, r); }
int tb_gotoh_pp(tb_ctx* ctx, const tb_batch* b, tb_score sc, tb_align_config ac, tb_result* r) { return run_gotoh(ctx, tb::kModePP, b, sc, ac, r); }
int tb_gotoh_ss(tb_ctx* ctx, const tb_batch* b, tb_score sc, tb_align_config ac, tb_result* r) { return run_gotoh(ctx, tb::kModeSS, b, sc, ac, r); }

int tb_rows_from_ops(int kind, const void* a1, int32_t len1, const void* a2, int32_t len2, const uint8_t* ops, int32_t L, char* row0, char* row1) {
  if (!ops || !row0 || !row1 || L < 0 || kind < 0 || kind > 2) return TB_ERR_INVALID;
  int r = 0, c = 0;
  for (int j = 0; j < L; ++j) {
    char x = '-', y = '-';
    if (ops[j] != 'h') {
      if (r >= len1) return TB_ERR_INVALID;
      x = kind == 1 ? ((const char*)a1)[r] : cons_char((const float*)a1, len1, r);
      ++r;
    }
    if (ops[j] != 'v') {
      if (c >= len2) return TB_ERR_INVALID;
      if (kind == 0) y = cons_char((const float*)a2, len2, c);
      else if (kind == 1) y = ((const char*)a2)[c];
      else {   // one-hot profile of a sequence (src/align.h:121-136) seen through _profileConsChar
        switch (((const char*)a2)[c]) {
          case 'A': case 'a': y = 'A'; break;
          case 'C': case 'c': y = 'C'; break;
          case 'G': case 'g': y = 'G'; break;
          case 'T': case 't': y = 'T'; break;
          case 'N': case 'n': case '-': y = 'N'; break;
          default: y = 'A'; break;   // all-zero column: no row exceeds p[0], index 0 wins
        }
      }
      ++c;
    }
    row0[j] = x; row1[j] = y;
  }
  return TB_OK;
}

int tb_unpack_ops(const uint8_t* packed, int32_t L, uint8_t* out) {
  if (!packed || !out || L < 0) return TB_ERR_INVALID;
  static const uint8_t sym[4] = {'s', 'h', 'v', '?'};
  for (int32_t j = 0; j < L; ++j) out[j] = sym[(packed[j >> 2] >> (2 * (j & 3))) & 3];
  return TB_OK;
}

int tb_decompose_sweep(tb_ctx* ctx, const tb_sweep_batch* batch, tb_sweep_result* res) {
  if (!ctx) return TB_ERR_INVALID;
  if (!batch || !res) return fail(ctx, TB_ERR_INVALID, "null batch/result");
  ctx->last_sweep_ms = 0;
  const size_t nt = batch->ntraces;
  if (nt == 0) return TB_OK;
  if (nt > (size_t)INT_MAX) return fail(ctx, TB_ERR_INVALID, "ntraces too large");
  if (!batch->refrow.base || !batch->refrow.off || !batch->refrow.len || !batch->primary.base || !batch->primary.off ||
      !batch->primary.len || !batch->secondary_base || !batch->vi_end || !batch->align_index || !batch->var_index ||
      !batch->ndel || !batch->nins || !res->fref || !res->fins || res->out_stride <= 0)
    return fail(ctx, TB_ERR_INVALID, "null pointer in sweep batch/result");
  TB_CUDA(ctx, cudaSetDevice(ctx->device));
  Lane& L = ctx->lanes[0];
  const int os = res->out_stride;
  tb::SweepBatch S{};
  S.out_stride = os;
  if (batch->mem == TB_MEM_DEVICE) {
    S.ref_base = (const char*)batch->refrow.base; S.ref_off = batch->refrow.off; S.ref_len = batch->refrow.len;
    S.pri_base = (const char*)batch->primary.base; S.sec_base = batch->secondary_base; S.bc_off = batch->primary.off;
    S.vi_end = batch->vi_end; S.align_index = batch->align_index; S.var_index = batch->var_index;
    S.ndel = batch->ndel; S.nins = batch->nins;
    S.fref = res->fref; S.fins = res->fins; S.grid = res->grid;
    TB_CUDA(ctx, cudaEventRecord(L.k0, L.stream));
    TB_CUDA(ctx, tb::launch_sweep(S, (int)nt, res->grid != nullptr, 2ll * os + (res->grid ? (long long)os * os : 0ll), L.stream));
    ctx->launches++;
    TB_CUDA(ctx, cudaEventRecord(L.k1, L.stream));
    TB_CUDA(ctx, cudaStreamSynchronize(L.stream));
    TB_CUDA(ctx, cudaEventElapsedTime(&ctx->last_sweep_ms, L.k0, L.k1));
    return TB_OK;
  }
  if (batch->mem != TB_MEM_HOST) return fail(ctx, TB_ERR_INVALID, "batch->mem must be TB_MEM_HOST or TB_MEM_DEVICE");
  // Host mode: strings are small (a few kB per trace); stage the whole batch in one go.
  long long rmin = LLONG_MAX, rmax = LLONG_MIN, pmin = LLONG_MAX, pmax = LLONG_MIN;
  for (size_t i = 0; i < nt; ++i) {
    if (batch->refrow.len[i] < 0 || batch->primary.len[i] < 0 || batch->refrow.off[i] < 0 || batch->primary.off[i] < 0)
      return fail(ctx, TB_ERR_INVALID, "negative offset/length");
    if (batch->ndel[i] < 0 || batch->nins[i] < 0 || batch->ndel[i] > os || batch->nins[i] > os)
      return fail(ctx, TB_ERR_INVALID, "ndel/nins outside [0, out_stride]");
    rmin = std::min<long long>(rmin, batch->refrow.off[i]); rmax = std::max<long long>(rmax, batch->refrow.off[i] + batch->refrow.len[i]);
    pmin = std::min<long long>(pmin, batch->primary.off[i]); pmax = std::max<long long>(pmax, batch->primary.off[i] + batch->primary.len[i]);
  }
  const size_t rbytes = (size_t)(rmax - rmin), pbytes = (size_t)(pmax - pmin);
  // layout of the single staging buffer (device): ref | pri | sec | ref_off | bc_off | 6 x int32[nt]
  auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t o_ref = 0, o_pri = al(o_ref + rbytes), o_sec = al(o_pri + pbytes), o_roff = al(o_sec + pbytes),
               o_boff = o_roff + nt * 8, o_i32 = o_boff + nt * 8, total = o_i32 + 6 * nt * 4;
  TB_CUDA(ctx, L.a.reserve(total));
  TB_CUDA(ctx, L.meta.reserve(nt * 16));
  const size_t out_elems = nt * (size_t)os;
  TB_CUDA(ctx, L.scores.reserve(2 * out_elems * 4));
  if (res->grid) TB_CUDA(ctx, L.ops.reserve(out_elems * (size_t)os * 4));
  char* d = L.a.as<char>();
  int64_t* hoff = static_cast<int64_t*>(L.meta.p);
  for (size_t i = 0; i < nt; ++i) { hoff[i] = batch->refrow.off[i] - rmin; hoff[nt + i] = batch->primary.off[i] - pmin; }
  cudaStream_t st = L.stream;
  TB_CUDA(ctx, cudaMemcpyAsync(d + o_ref, (const char*)batch->refrow.base + rmin, rbytes, cudaMemcpyHostToDevice, st));
  TB_CUDA(ctx, cudaMemcpyAsync(d + o_pri, (const char*)batch->primary.base + pmin, pbytes, cudaMemcpyHostToDevice, st));
  TB_CUDA(ctx, cudaMemcpyAsync(d + o_sec, batch->secondary_base + pmin, pbytes, cudaMemcpyHostToDevice, st));
  TB_CUDA(ctx, cudaMemcpyAsync(d + o_roff, hoff, nt * 8, cudaMemcpyHostToDevice, st));
  TB_CUDA(ctx, cudaMemcpyAsync(d + o_boff, hoff + nt, nt * 8, cudaMemcpyHostToDevice, st));
  const int32_t* i32src[6] = {batch->refrow.len, batch->vi_end, batch->align_index, batch->var_index, batch->ndel, batch->nins};
  for (int k = 0; k < 6; ++k) TB_CUDA(ctx, cudaMemcpyAsync(d + o_i32 + (size_t)k * nt * 4, i32src[k], nt * 4, cudaMemcpyHostToDevice, st));
  ctx->h2d += rbytes + 2 * pbytes + nt * 16 + 6 * nt * 4;
  S.ref_base = d + o_ref; S.pri_base = d + o_pri; S.sec_base = d + o_sec;
  S.ref_off = (const int64_t*)(d + o_roff); S.bc_off = (const int64_t*)(d + o_boff);
  const int32_t* i32 = (const int32_t*)(d + o_i32);
  S.ref_len = i32; S.vi_end = i32 + nt; S.align_index = i32 + 2 * nt; S.var_index = i32 + 3 * nt; S.ndel = i32 + 4 * nt; S.nins = i32 + 5 * nt;
  S.fref = L.scores.as<int32_t>(); S.fins = S.fref + out_elems; S.grid = res->grid ? L.ops.as<int32_t>() : nullptr;
  TB_CUDA(ctx, cudaEventRecord(L.k0, st));
  long long max_tasks = 1;
  for (size_t i = 0; i < nt; ++i)
    max_tasks = std::max(max_tasks, (long long)batch->ndel[i] + batch->nins[i] + (res->grid ? (long long)batch->ndel[i] * batch->nins[i] : 0ll));
  TB_CUDA(ctx, tb::launch_sweep(S, (int)nt, res->grid != nullptr, max_tasks, st));
  ctx->launches++;
  TB_CUDA(ctx, cudaEventRecord(L.k1, st));
  TB_CUDA(ctx, cudaMemcpyAsync(res->fref, S.fref, out_elems * 4, cudaMemcpyDeviceToHost, st));
  TB_CUDA(ctx, cudaMemcpyAsync(res->fins, S.fins, out_elems * 4, cudaMemcpyDeviceToHost, st));
  if (res->grid) TB_CUDA(ctx, cudaMemcpyAsync(res->grid, S.grid, out_elems * (size_t)os * 4, cudaMemcpyDeviceToHost, st));
  ctx->d2h += 2 * out_elems * 4 + (res->grid ? out_elems * (size_t)os * 4 : 0);
  TB_CUDA(ctx, cudaStreamSynchronize(st));
  TB_CUDA(ctx, cudaEventElapsedTime(&ctx->last_sweep_ms, L.k0, L.k1));
  return TB_OK;
}

// ---- profile construction ---------------------------------------------------------------------------------------
namespace {
// Stream-ordered staging of one host array (freed by the caller after the stream drains).
struct Staged {
  std::vector<void*> bufs;
  cudaStream_t st;
  explicit Staged(cudaStream_t s) : st(s) {}
  ~Staged() { for (void* p : bufs) cudaFreeAsync(p, st); }
  cudaError_t alloc(void** out, size_t bytes) {
    cudaError_t e = cudaMallocAsync(out, bytes ? bytes : 1, st);
    if (e == cudaSuccess) bufs.push_back(*out);
    return e;
  }
  cudaError_t up(void** out, const void* src, size_t bytes) {
    cudaError_t e = alloc(out, bytes);
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(*out, src, bytes, cudaMemcpyHostToDevice, st);
  }
};
}  // namespace

// ---- device-resident trace samples ---------------------------------------------------------------------------------
struct tb_trace_set {
  int device = 0;
  size_t n = 0;
  int32_t* d_base = nullptr; int64_t* d_off = nullptr; int32_t* d_len = nullptr;
  std::vector<int64_t> off; std::vector<int32_t> len;      // host copies (validation, subsets)
  uint64_t bytes = 0;
};

int tb_trace_set_destroy(tb_ctx* ctx, tb_trace_set* set) {
  if (!set) return TB_OK;
  if (ctx) cudaSetDevice(ctx->device);
  if (set->d_base) cudaFree(set->d_base);
  if (set->d_off) cudaFree(set->d_off);
  if (set->d_len) cudaFree(set->d_len);
  delete set;
  return TB_OK;
}

int tb_trace_set_info(const tb_trace_set* set, size_t* ntraces, uint64_t* device_bytes, tb_arena* device_arena) {
  if (!set) return TB_ERR_INVALID;
  if (ntraces) *ntraces = set->n;
  if (device_bytes) *device_bytes = set->bytes;
  if (device_arena) { device_arena->base = set->d_base; device_arena->off = set->d_off; device_arena->len = set->d_len; }
  return TB_OK;
}

int tb_trace_set_create(tb_ctx* ctx, const int32_t* const* channels, const int32_t* nsamples, size_t nt, tb_trace_set** out) {
  if (!ctx) return TB_ERR_INVALID;
  if (!out || (nt && (!channels || !nsamples))) return fail(ctx, TB_ERR_INVALID, "null argument");
  if (nt > (size_t)INT_MAX) return fail(ctx, TB_ERR_INVALID, "ntraces too large");
  TB_CUDA(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<tb_trace_set> S(new tb_trace_set);
  S->device = ctx->device; S->n = nt; S->off.resize(nt); S->len.resize(nt);
  int64_t total = 0, biggest = 0;
  for (size_t t = 0; t < nt; ++t) {
    if (nsamples[t] < 0) return fail(ctx, TB_ERR_INVALID, "negative trace length");
    for (int k = 0; k < 4; ++k) if (nsamples[t] && !channels[4 * t + k]) return fail(ctx, TB_ERR_INVALID, "null channel pointer");
    S->off[t] = total; S->len[t] = nsamples[t];
    total += 4ll * nsamples[t]; biggest = std::max<int64_t>(biggest, 4ll * nsamples[t]);
  }
  S->bytes = (uint64_t)std::max<int64_t>(total, 1) * 4 + nt * 12;
  cudaStream_t st = ctx->lanes[0].stream;
  auto cleanup = [&](cudaError_t e, const char* what) { tb_trace_set* raw = S.release(); tb_trace_set_destroy(ctx, raw); return cuda_fail(ctx, e, what); };
  cudaError_t e;
  if ((e = cudaMalloc((void**)&S->d_base, (size_t)std::max<int64_t>(total, 1) * 4)) != cudaSuccess) return cleanup(e, "cudaMalloc(trace samples)");
  if ((e = cudaMalloc((void**)&S->d_off, std::max<size_t>(nt, 1) * 8)) != cudaSuccess) return cleanup(e, "cudaMalloc(trace offsets)");
  if ((e = cudaMalloc((void**)&S->d_len, std::max<size_t>(nt, 1) * 4)) != cudaSuccess) return cleanup(e, "cudaMalloc(trace lengths)");
  if (nt) {
    if ((e = cudaMemcpyAsync(S->d_off, S->off.data(), nt * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return cleanup(e, "cudaMemcpyAsync(offsets)");
    if ((e = cudaMemcpyAsync(S->d_len, S->len.data(), nt * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return cleanup(e, "cudaMemcpyAsync(lengths)");
  }
  // two pinned staging buffers: host threads gather the scattered channel vectors into one while the other crosses PCIe
  const int64_t chunk_elems = std::max<int64_t>(biggest, (int64_t)8 << 20);      // >= 32 MB, and never less than one trace
  PinBuf pin[2];
  cudaEvent_t done[2] = {nullptr, nullptr};
  auto drop = [&]() { for (int i = 0; i < 2; ++i) { pin[i].release(); if (done[i]) cudaEventDestroy(done[i]); } };
  for (int i = 0; i < 2 && total; ++i) {
    if ((e = pin[i].reserve((size_t)std::min<int64_t>(chunk_elems, total) * 4)) != cudaSuccess) { drop(); return cleanup(e, "cudaHostAlloc(trace staging)"); }
    if ((e = cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming)) != cudaSuccess) { drop(); return cleanup(e, "cudaEventCreate"); }
  }
  const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  size_t t0 = 0;
  int which = 0;
  bool used[2] = {false, false};
  while (t0 < nt) {
    size_t t1 = t0;
    int64_t elems = 0;
    while (t1 < nt && elems + 4ll * nsamples[t1] <= chunk_elems) { elems += 4ll * nsamples[t1]; ++t1; }
    if (elems) {
      if (used[which] && (e = cudaEventSynchronize(done[which])) != cudaSuccess) { drop(); return cleanup(e, "cudaEventSynchronize"); }
      int32_t* dst = static_cast<int32_t*>(pin[which].p);
      const int64_t base_off = S->off[t0];
      const unsigned nthr = (unsigned)std::min<size_t>(hw, t1 - t0);
      std::vector<std::thread> pool;
      for (unsigned w = 0; w < nthr; ++w)
        pool.emplace_back([&, w]() {
          for (size_t t = t0 + w; t < t1; t += nthr)
            for (int k = 0; k < 4; ++k)
              if (nsamples[t]) std::memcpy(dst + (S->off[t] - base_off) + (int64_t)k * nsamples[t], channels[4 * t + k], (size_t)nsamples[t] * 4);
        });
      for (auto& th : pool) th.join();
      if ((e = cudaMemcpyAsync(S->d_base + base_off, dst, (size_t)elems * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) { drop(); return cleanup(e, "cudaMemcpyAsync(samples)"); }
      if ((e = cudaEventRecord(done[which], st)) != cudaSuccess) { drop(); return cleanup(e, "cudaEventRecord"); }
      used[which] = true;
      which ^= 1;
    }
    t0 = t1;
  }
  e = cudaStreamSynchronize(st);
  drop();
  if (e != cudaSuccess) return cleanup(e, "cudaStreamSynchronize");
  ctx->h2d += (uint64_t)total * 4 + nt * 12;
  *out = S.release();
  return TB_OK;
}

namespace {
// The trace arena of a basecall / profile / fraction batch on the device: the caller's host arena uploaded (plain TB_MEM_HOST) or
// a view into a tb_trace_set (TB_TRACE_SET). host_off / host_len: what validation and sizing read.
struct TraceView {
  const int32_t* base = nullptr; const int64_t* off = nullptr; const int32_t* len = nullptr;   // device
  const int32_t* host_len = nullptr;
  std::vector<int64_t> sub_off; std::vector<int32_t> sub_len;
  uint64_t uploaded = 0;
};
int resolve_traces(tb_ctx* ctx, const tb_arena& a, int32_t mem, size_t nt, Staged& S, TraceView& V) {
  if (mem & TB_TRACE_SET) {
    const tb_trace_set* set = static_cast<const tb_trace_set*>(a.base);
    if (!set) return fail(ctx, TB_ERR_INVALID, "TB_TRACE_SET without a trace set in trace.base");
    if (set->device != ctx->device) return fail(ctx, TB_ERR_INVALID, "trace set lives on another device");
    V.base = set->d_base;
    if (!a.off) {
      if (nt != set->n) return fail(ctx, TB_ERR_INVALID, "ntraces differs from the trace set's size (pass indices in trace.off for a subset)");
      V.off = set->d_off; V.len = set->d_len; V.host_len = set->len.data();
      return TB_OK;
    }
    V.sub_off.resize(nt); V.sub_len.resize(nt);
    for (size_t i = 0; i < nt; ++i) {
      if (a.off[i] < 0 || (size_t)a.off[i] >= set->n) return fail(ctx, TB_ERR_INVALID, "trace index outside the trace set");
      V.sub_off[i] = set->off[(size_t)a.off[i]]; V.sub_len[i] = set->len[(size_t)a.off[i]];
    }
    void *d_o, *d_l;
    TB_CUDA(ctx, S.up(&d_o, V.sub_off.data(), nt * 8)); TB_CUDA(ctx, S.up(&d_l, V.sub_len.data(), nt * 4));
    V.off = (const int64_t*)d_o; V.len = (const int32_t*)d_l; V.host_len = V.sub_len.data();
    V.uploaded = nt * 12;
    return TB_OK;
  }
  if (!a.base || !a.off || !a.len) return fail(ctx, TB_ERR_INVALID, "null pointer in the trace arena");
  long long tmax = 0;
  for (size_t i = 0; i < nt; ++i) {
    if (a.len[i] < 0 || a.off[i] < 0) return fail(ctx, TB_ERR_INVALID, "negative trace offset/length");
    tmax = std::max<long long>(tmax, a.off[i] + 4ll * a.len[i]);
  }
  void *d_tr, *d_o, *d_l;
  TB_CUDA(ctx, S.up(&d_tr, a.base, (size_t)std::max(tmax, 1ll) * 4));
  TB_CUDA(ctx, S.up(&d_o, a.off, nt * 8)); TB_CUDA(ctx, S.up(&d_l, a.len, nt * 4));
  V.base = (const int32_t*)d_tr; V.off = (const int64_t*)d_o; V.len = (const int32_t*)d_l; V.host_len = a.len;
  V.uploaded = (uint64_t)tmax * 4 + nt * 12;
  return TB_OK;
}
// Outputs of a batch back to the host: one copy when the items tile the arena without holes, one per item otherwise.
extern "C++" template <typename T>
cudaError_t copy_items_back(T* host, const T* dev, const int64_t* off, const int32_t* cap, const int32_t* used, int mul, size_t nt, cudaStream_t st) {
  bool dense = true;
  int64_t at = nt ? off[0] : 0;
  for (size_t i = 0; i < nt && dense; ++i) { dense = off[i] == at; at += (int64_t)mul * cap[i]; }
  if (dense && nt) return cudaMemcpyAsync(host + off[0], dev + off[0], (size_t)(at - off[0]) * sizeof(T), cudaMemcpyDeviceToHost, st);
  for (size_t i = 0; i < nt; ++i) {
    cudaError_t e = cudaMemcpyAsync(host + off[i], dev + off[i], (size_t)mul * used[i] * sizeof(T), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}
}  // namespace

int tb_create_profile(tb_ctx* ctx, const tb_profile_batch* b, float* out_base, const int64_t* out_off, int32_t* out_len) {
  if (!ctx) return TB_ERR_INVALID;
  if (!b || !out_base || !out_off || !out_len) return fail(ctx, TB_ERR_INVALID, "null batch/output");
  const size_t nt = b->ntraces;
  if (nt == 0) return TB_OK;
  if (nt > (size_t)INT_MAX) return fail(ctx, TB_ERR_INVALID, "ntraces too large");
  const int32_t mem = b->mem & ~TB_TRACE_SET;
  if (!b->trace.base || (!(b->mem & TB_TRACE_SET) && (!b->trace.off || !b->trace.len)) || !b->bcpos.base || !b->bcpos.off || !b->bcpos.len || !b->primary_base ||
      !b->secondary_base)
    return fail(ctx, TB_ERR_INVALID, "null pointer in profile batch");
  TB_CUDA(ctx, cudaSetDevice(ctx->device));
  Lane& L = ctx->lanes[0];
  cudaStream_t st = L.stream;
  tb::ProfileBatch P{};
  if (b->mem == TB_MEM_DEVICE) {
    P.trace_base = (const int32_t*)b->trace.base; P.trace_off = b->trace.off; P.trace_len = b->trace.len;
    P.bcpos_base = (const int32_t*)b->bcpos.base; P.pri_base = b->primary_base; P.sec_base = b->secondary_base;
    P.bc_off = b->bcpos.off; P.bc_len = b->bcpos.len; P.trim_left = b->trim_left; P.trim_right = b->trim_right;
    P.out_base = out_base; P.out_off = out_off; P.out_len = out_len;
    TB_CUDA(ctx, tb::launch_create_profile(P, (int)nt, st));
    ctx->launches++;
    TB_CUDA(ctx, cudaStreamSynchronize(st));
    return TB_OK;
  }
  if (mem != TB_MEM_HOST) return fail(ctx, TB_ERR_INVALID, "batch->mem must be TB_MEM_HOST (optionally with TB_TRACE_SET) or TB_MEM_DEVICE");
  long long bmax = 0, omax = 0;
  for (size_t i = 0; i < nt; ++i) {
    if (b->bcpos.len[i] < 0 || b->bcpos.off[i] < 0 || out_off[i] < 0) return fail(ctx, TB_ERR_INVALID, "negative offset/length");
    bmax = std::max<long long>(bmax, b->bcpos.off[i] + b->bcpos.len[i]);
    omax = std::max<long long>(omax, out_off[i] + 6ll * b->bcpos.len[i]);
  }
  {
    Staged S(st);
    TraceView V;
    if (int rc = resolve_traces(ctx, b->trace, b->mem, nt, S, V)) return rc;
    const void *d_tr = V.base, *d_toff = V.off, *d_tlen = V.len;
    const long long tmax = (long long)(V.uploaded / 4);
    void *d_bp, *d_pri, *d_sec, *d_boff, *d_blen, *d_tl = nullptr, *d_trr = nullptr, *d_out, *d_ooff, *d_olen;
    TB_CUDA(ctx, S.up(&d_bp, b->bcpos.base, (size_t)bmax * 4));
    TB_CUDA(ctx, S.up(&d_pri, b->primary_base, (size_t)bmax)); TB_CUDA(ctx, S.up(&d_sec, b->secondary_base, (size_t)bmax));
    TB_CUDA(ctx, S.up(&d_boff, b->bcpos.off, nt * 8)); TB_CUDA(ctx, S.up(&d_blen, b->bcpos.len, nt * 4));
    if (b->trim_left) TB_CUDA(ctx, S.up(&d_tl, b->trim_left, nt * 4));
    if (b->trim_right) TB_CUDA(ctx, S.up(&d_trr, b->trim_right, nt * 4));
    TB_CUDA(ctx, S.up(&d_ooff, out_off, nt * 8));
    TB_CUDA(ctx, S.alloc(&d_out, (size_t)omax * 4)); TB_CUDA(ctx, S.alloc(&d_olen, nt * 4));
    ctx->h2d += (size_t)tmax * 4 + (size_t)bmax * 6 + nt * 32;
    P.trace_base = (const int32_t*)d_tr; P.trace_off = (const int64_t*)d_toff; P.trace_len = (const int32_t*)d_tlen;
    P.bcpos_base = (const int32_t*)d_bp; P.pri_base = (const char*)d_pri; P.sec_base = (const char*)d_sec;
    P.bc_off = (const int64_t*)d_boff; P.bc_len = (const int32_t*)d_blen; P.trim_left = (const int32_t*)d_tl; P.trim_right = (const int32_t*)d_trr;
    P.out_base = (float*)d_out; P.out_off = (const int64_t*)d_ooff; P.out_len = (int32_t*)d_olen;
    TB_CUDA(ctx, tb::launch_create_profile(P, (int)nt, st));
    ctx->launches++;
    TB_CUDA(ctx, cudaMemcpyAsync(out_len, d_olen, nt * 4, cudaMemcpyDeviceToHost, st));
    TB_CUDA(ctx, cudaStreamSynchronize(st));
    // items may be sparse in the caller's output arena: one copy when they tile it, else each item's 6 * sz floats
    TB_CUDA(ctx, copy_items_back(out_base, (const float*)d_out, out_off, b->bcpos.len, out_len, 6, nt, st));
    ctx->d2h += (size_t)omax * 4 + nt * 4;
    TB_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return TB_OK;
}

int tb_basecall(tb_ctx* ctx, const tb_basecall_batch* b, float sigratio, int32_t* bcpos_out, char* primary_out, char* secondary_out,
                char* consensus_out, const int64_t* out_off, int32_t* out_len) {
  if (!ctx) return TB_ERR_INVALID;
  if (!b || !bcpos_out || !primary_out || !secondary_out || !consensus_out || !out_off || !out_len) return fail(ctx, TB_ERR_INVALID, "null batch/output");
  const size_t nt = b->ntraces;
  if (nt == 0) return TB_OK;
  if (nt > (size_t)INT_MAX) return fail(ctx, TB_ERR_INVALID, "ntraces too large");
  if (!b->trace.base || (!(b->mem & TB_TRACE_SET) && (!b->trace.off || !b->trace.len)) || !b->ploc.base || !b->ploc.off || !b->ploc.len)
    return fail(ctx, TB_ERR_INVALID, "null arena pointer");
  TB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->lanes[0].stream;
  tb::BasecallBatch P{};
  P.sigratio = sigratio;
  if (b->mem == TB_MEM_DEVICE) {
    P.trace_base = (const int32_t*)b->trace.base; P.trace_off = b->trace.off; P.trace_len = b->trace.len;
    P.ploc_base = (const int32_t*)b->ploc.base; P.ploc_off = b->ploc.off; P.ploc_len = b->ploc.len;
    P.bcpos_out = bcpos_out; P.pri_out = primary_out; P.sec_out = secondary_out; P.con_out = consensus_out; P.out_off = out_off; P.out_len = out_len;
    TB_CUDA(ctx, tb::launch_basecall(P, (int)nt, st));
    ctx->launches++;
    TB_CUDA(ctx, cudaStreamSynchronize(st));
    return TB_OK;
  }
  if ((b->mem & ~TB_TRACE_SET) != TB_MEM_HOST) return fail(ctx, TB_ERR_INVALID, "batch->mem must be TB_MEM_HOST (optionally with TB_TRACE_SET) or TB_MEM_DEVICE");
  {
    Staged S(st);
    TraceView V;
    if (int rc = resolve_traces(ctx, b->trace, b->mem, nt, S, V)) return rc;
    long long pmax = 0, omax = 0;
    for (size_t i = 0; i < nt; ++i) {
      if (V.host_len[i] < 2 || b->ploc.len[i] < 0 || b->ploc.off[i] < 0 || out_off[i] < 0)
        return fail(ctx, TB_ERR_INVALID, "negative offset/length (a trace needs at least 2 samples)");
      for (int32_t j = 0; j < b->ploc.len[i]; ++j) {
        const int32_t v = ((const int32_t*)b->ploc.base)[b->ploc.off[i] + j];
        if (v < 0 || v >= V.host_len[i]) return fail(ctx, TB_ERR_INVALID, "basecall position outside the trace");
      }
      pmax = std::max<long long>(pmax, b->ploc.off[i] + b->ploc.len[i]);
      omax = std::max<long long>(omax, out_off[i] + b->ploc.len[i]);
    }
    const void *d_tr = V.base, *d_toff = V.off, *d_tlen = V.len;
    const long long tmax = (long long)(V.uploaded / 4);
    void *d_pl, *d_poff, *d_plen, *d_ooff, *d_olen, *d_pos, *d_pri, *d_sec, *d_con;
    TB_CUDA(ctx, S.up(&d_pl, b->ploc.base, (size_t)pmax * 4));
    TB_CUDA(ctx, S.up(&d_poff, b->ploc.off, nt * 8)); TB_CUDA(ctx, S.up(&d_plen, b->ploc.len, nt * 4));
    TB_CUDA(ctx, S.up(&d_ooff, out_off, nt * 8));
    TB_CUDA(ctx, S.alloc(&d_olen, nt * 4)); TB_CUDA(ctx, S.alloc(&d_pos, (size_t)omax * 4));
    TB_CUDA(ctx, S.alloc(&d_pri, (size_t)omax)); TB_CUDA(ctx, S.alloc(&d_sec, (size_t)omax)); TB_CUDA(ctx, S.alloc(&d_con, (size_t)omax));
    P.trace_base = (const int32_t*)d_tr; P.trace_off = (const int64_t*)d_toff; P.trace_len = (const int32_t*)d_tlen;
    P.ploc_base = (const int32_t*)d_pl; P.ploc_off = (const int64_t*)d_poff; P.ploc_len = (const int32_t*)d_plen;
    P.bcpos_out = (int32_t*)d_pos; P.pri_out = (char*)d_pri; P.sec_out = (char*)d_sec; P.con_out = (char*)d_con;
    P.out_off = (const int64_t*)d_ooff; P.out_len = (int32_t*)d_olen;
    TB_CUDA(ctx, tb::launch_basecall(P, (int)nt, st));
    ctx->launches++;
    TB_CUDA(ctx, cudaMemcpyAsync(out_len, d_olen, nt * 4, cudaMemcpyDeviceToHost, st));
    TB_CUDA(ctx, cudaStreamSynchronize(st));
    TB_CUDA(ctx, copy_items_back(bcpos_out, (const int32_t*)d_pos, out_off, b->ploc.len, out_len, 1, nt, st));
    TB_CUDA(ctx, copy_items_back(primary_out, (const char*)d_pri, out_off, b->ploc.len, out_len, 1, nt, st));
    TB_CUDA(ctx, copy_items_back(secondary_out, (const char*)d_sec, out_off, b->ploc.len, out_len, 1, nt, st));
    TB_CUDA(ctx, copy_items_back(consensus_out, (const char*)d_con, out_off, b->ploc.len, out_len, 1, nt, st));
    ctx->h2d += (size_t)tmax * 4 + (size_t)pmax * 4 + nt * 32; ctx->d2h += (size_t)omax * 7 + nt * 4;
    TB_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return TB_OK;
}

int tb_revcomp_profile(tb_ctx* ctx, const tb_arena* in, size_t n, int32_t mem, float* out_base, const int64_t* out_off) {
  if (!ctx) return TB_ERR_INVALID;
  if (!in || !out_base || !out_off) return fail(ctx, TB_ERR_INVALID, "null input/output");
  if (n == 0) return TB_OK;
  if (n > (size_t)INT_MAX) return fail(ctx, TB_ERR_INVALID, "too many profiles");
  if (!in->base || !in->off || !in->len) return fail(ctx, TB_ERR_INVALID, "null arena pointer");
  TB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->lanes[0].stream;
  if (mem == TB_MEM_DEVICE) {
    TB_CUDA(ctx, tb::launch_revcomp_profile((const float*)in->base, in->off, in->len, out_base, out_off, (int)n, st));
    ctx->launches++;
    TB_CUDA(ctx, cudaStreamSynchronize(st));
    return TB_OK;
  }
  if (mem != TB_MEM_HOST) return fail(ctx, TB_ERR_INVALID, "mem must be TB_MEM_HOST or TB_MEM_DEVICE");
  long long imax = 0, omax = 0;
  for (size_t i = 0; i < n; ++i) {
    if (in->len[i] < 0 || in->off[i] < 0 || out_off[i] < 0) return fail(ctx, TB_ERR_INVALID, "negative offset/length");
    imax = std::max<long long>(imax, in->off[i] + 6ll * in->len[i]);
    omax = std::max<long long>(omax, out_off[i] + 6ll * in->len[i]);
  }
  {
    Staged S(st);
    void *d_in, *d_ioff, *d_len, *d_ooff, *d_out;
    TB_CUDA(ctx, S.up(&d_in, in->base, (size_t)imax * 4)); TB_CUDA(ctx, S.up(&d_ioff, in->off, n * 8));
    TB_CUDA(ctx, S.up(&d_len, in->len, n * 4)); TB_CUDA(ctx, S.up(&d_ooff, out_off, n * 8));
    TB_CUDA(ctx, S.alloc(&d_out, (size_t)omax * 4));
    TB_CUDA(ctx, tb::launch_revcomp_profile((const float*)d_in, (const int64_t*)d_ioff, (const int32_t*)d_len, (float*)d_out, (const int64_t*)d_ooff, (int)n, st));
    ctx->launches++;
    for (size_t i = 0; i < n; ++i)
      TB_CUDA(ctx, cudaMemcpyAsync(out_base + out_off[i], (const float*)d_out + out_off[i], (size_t)6 * in->len[i] * 4, cudaMemcpyDeviceToHost, st));
    ctx->h2d += (size_t)imax * 4 + n * 20; ctx->d2h += (size_t)omax * 4;
    TB_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return TB_OK;
}

int tb_trim_reference_slice(const char* row0, const char* row1, int32_t L, int32_t refslice_len, int32_t forward, uint32_t pos,
                            int32_t trim_left, int32_t trim_right, int32_t* ri_out, int32_t* risize_out, uint32_t* new_pos) {
  if (!row0 || !row1 || L < 0 || refslice_len < 0 || !ri_out || !risize_out || !new_pos) return TB_ERR_INVALID;
  // reference src/fmindex.h:432-446: ri = reference characters before the first trace column; [s, e) = trace extent
  uint32_t ri = 0;
  int32_t s = -1, e = -1;
  for (int32_t j = 0; j < L; ++j) {
    if (row0[j] != '-') { if (s == -1) s = j; e = j + 1; }
    if (s == -1 && row1[j] != '-') ++ri;
  }
  uint32_t risize = 0;
  for (int32_t j = s; j < e; ++j) if (row1[j] != '-') ++risize;
  const uint32_t tl = (uint16_t)trim_left, tr = (uint16_t)trim_right;       // config fields are uint16_t (src/sage.h:37-56)
  if (ri >= tl) { ri -= tl; risize += tl; }                                  // src/fmindex.h:447-450
  if (ri + risize + tr < (uint32_t)refslice_len) risize += tr;               // src/fmindex.h:451
  // std::string::substr(ri, risize) clamps the length to the end of the slice (and throws when ri > size)
  if (ri > (uint32_t)refslice_len) return TB_ERR_INVALID;
  const uint32_t eff = std::min<uint32_t>(risize, (uint32_t)refslice_len - ri);
  *ri_out = (int32_t)ri; *risize_out = (int32_t)eff;
  if (forward) *new_pos = pos + ri;                                          // src/fmindex.h:454
  else {
    const int32_t offset = refslice_len - (int32_t)ri - (int32_t)risize;    // uses the UNclamped risize, as the reference does
    *new_pos = offset < 0 ? pos : pos + (uint32_t)offset;                    // src/fmindex.h:455-462 (warning text not reproduced)
  }
  return TB_OK;
}

int tb_find_breakpoint(const float* p, int32_t len, int32_t* indelshift, int32_t* traceleft, uint32_t* breakpoint, float* best_diff) {
  if (!p || len < 0 || !indelshift || !traceleft || !breakpoint || !best_diff) return TB_ERR_INVALID;
  std::vector<double> sig((size_t)len);
  for (int32_t j = 0; j < len; ++j) {                                        // src/decompose.h:11-24
    double best = 0.001, snd = 0.001;
    for (int i = 0; i < 6; ++i) {
      const float v = p[(size_t)i * len + j];
      if (v > best) { snd = best; best = v; }
      else if (v > snd) snd = v;
    }
    sig[j] = best - snd;
  }
  float bestDiff = 0;                                                        // TraceBreakpoint::bestDiff is a float (src/fmindex.h:55)
  bool left_flag = true;
  uint32_t bp = 0;
  const uint32_t w = 25, n = (uint32_t)len;
  if (w < n) {
    for (uint32_t i = w; i < n - w; ++i) {                                   // src/decompose.h:31-47
      double ls = 0; for (uint32_t k = i - w; k < i; ++k) ls += sig[k];
      const double left = ls / (double)w;
      double rs = 0; for (uint32_t k = i; k < i + w; ++k) rs += sig[k];
      const double right = rs / (double)w;
      const double diff = std::abs(right - left);
      if (diff > bestDiff) { bp = i; bestDiff = (float)diff; left_flag = !(left < right); }
    }
  }
  int shift = 1;
  if (bestDiff < 0.25) { shift = 0; bp = n; left_flag = true; bestDiff = 0; } // src/decompose.h:49-55
  *indelshift = shift; *traceleft = left_flag ? 1 : 0; *breakpoint = bp; *best_diff = bestDiff;
  return TB_OK;
}

}  // extern "C"

// ---- reference anchoring (kernels in anchor.cu) -----------------------------------------
struct tb_index {
  int device = 0;
  long long n = 0;
  unsigned char* text = nullptr;
  uint4* rec = nullptr;
  uint2* dir = nullptr;
  int dir_chars = 0;
  size_t bytes = 0;
};

int tb_index_destroy(tb_ctx* ctx, tb_index* idx) {
  if (!idx) return TB_OK;
  if (ctx) cudaSetDevice(ctx->device);
  cudaFree(idx->text); cudaFree(idx->rec); cudaFree(idx->dir);
  delete idx;
  return TB_OK;
}

int tb_index_build(tb_ctx* ctx, const char* text, int64_t text_len, int32_t mem, tb_index** out) {
  if (!ctx) return TB_ERR_INVALID;
  if (!text || !out || text_len <= 0) return fail(ctx, TB_ERR_INVALID, "null/empty text");
  if (text_len >= (int64_t)UINT_MAX) return fail(ctx, TB_ERR_UNSUPPORTED, "text longer than 2^32-2 characters");
  if (mem != TB_MEM_HOST && mem != TB_MEM_DEVICE) return fail(ctx, TB_ERR_INVALID, "mem must be TB_MEM_HOST or TB_MEM_DEVICE");
  TB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->lanes[0].stream;
  const long long n = text_len;
  tb_index* X = new (std::nothrow) tb_index();
  if (!X) return fail(ctx, TB_ERR_NOMEM, "out of host memory");
  X->device = ctx->device; X->n = n;
  unsigned long long *keys_a = nullptr, *keys_b = nullptr; unsigned *pos_a = nullptr, *pos_b = nullptr; void* temp = nullptr; int* d_invalid = nullptr;
  size_t temp_bytes = 0;
  X->dir_chars = tb::index_dir_chars(n);
  const size_t nd = (size_t)1 << (2 * X->dir_chars);
  auto cleanup = [&](int rc) {
    cudaFree(keys_a); cudaFree(pos_a); cudaFree(keys_b); cudaFree(pos_b); cudaFree(temp); cudaFree(d_invalid);
    if (rc != TB_OK) tb_index_destroy(ctx, X);
    return rc;
  };
#define TB_IDX(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cleanup(cuda_fail(ctx, e__, #call)); } while (0)
  TB_IDX(tb::index_sort_temp_bytes(n, &temp_bytes));
  TB_IDX(cudaMalloc(&X->text, (size_t)n + 64));
  TB_IDX(cudaMalloc(&keys_a, (size_t)n * 8)); TB_IDX(cudaMalloc(&pos_a, (size_t)n * 4));
  TB_IDX(cudaMalloc(&keys_b, (size_t)n * 8)); TB_IDX(cudaMalloc(&pos_b, (size_t)n * 4));
  TB_IDX(cudaMalloc(&temp, temp_bytes ? temp_bytes : 1)); TB_IDX(cudaMalloc(&d_invalid, 4));
  TB_IDX(cudaMemcpyAsync(X->text, text, (size_t)n, mem == TB_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
  if (mem == TB_MEM_HOST) ctx->h2d += (size_t)n;
  TB_IDX(tb::index_sort_text(X->text, n, keys_a, pos_a, keys_b, pos_b, temp, temp_bytes, d_invalid, st));
  int invalid = 0;
  TB_IDX(cudaMemcpyAsync(&invalid, d_invalid, 4, cudaMemcpyDeviceToHost, st));
  TB_IDX(cudaStreamSynchronize(st));
  // the sort's input columns and scratch go before the records come: 28 B per character at the peak instead of 40
  cudaFree(keys_a); keys_a = nullptr; cudaFree(pos_a); pos_a = nullptr; cudaFree(temp); temp = nullptr;
  if (!invalid) {
    TB_IDX(cudaMalloc(&X->rec, (size_t)n * 16)); TB_IDX(cudaMalloc(&X->dir, nd * 8));
    TB_IDX(tb::index_make_records(keys_b, pos_b, n, X->rec, st));
    TB_IDX(tb::index_make_dir(X->rec, n, X->dir, X->dir_chars, st));
    TB_IDX(cudaStreamSynchronize(st));
  }
  ctx->launches += 4;
#undef TB_IDX
  if (invalid) return cleanup(fail(ctx, TB_ERR_UNSUPPORTED, "text holds a byte outside ACGTN, the IUPAC codes RYSWKMBDHV and '\\n'"));
  X->bytes = (size_t)n * 17 + nd * 8;
  *out = X;
  return cleanup(TB_OK);
}

int tb_index_info(const tb_index* idx, int64_t* text_len, uint64_t* device_bytes, const char** device_text) {
  if (!idx) return TB_ERR_INVALID;
  if (text_len) *text_len = idx->n;
  if (device_bytes) *device_bytes = idx->bytes;
  if (device_text) *device_text = reinterpret_cast<const char*>(idx->text);
  return TB_OK;
}

namespace {
unsigned next_pow2(unsigned long long v) { unsigned long long p = 2; while (p < v) p <<= 1; return (unsigned)std::min<unsigned long long>(p, 1ull << 31); }

// Global-table path over the traces in `todo` (host list): count -> size the tables -> fill -> decide. Updates the
// device result arrays in A; the caller re-reads `pass`.
int anchor_global_pass(tb_ctx* ctx, const tb::KmerIndexView& X, tb::AnchorBatch A, const std::vector<int32_t>& todo, bool nonunique, cudaStream_t st) {
  const size_t budget_elems = (size_t)1 << 27;            // 128 Mi slots (1.5 GB of keys + counts) per sub-batch
  size_t done = 0;
  while (done < todo.size()) {
    Staged S(st);
    // count on a window of traces, then cut it where the tables exceed the budget
    const size_t wn = std::min<size_t>(todo.size() - done, 65536);
    void *d_todo, *d_tot;
    TB_CUDA(ctx, S.up(&d_todo, todo.data() + done, wn * 4));
    TB_CUDA(ctx, S.alloc(&d_tot, wn * 16));
    TB_CUDA(ctx, cudaMemsetAsync(d_tot, 0, wn * 16, st));
    A.todo = (const int32_t*)d_todo; A.totals = (unsigned long long*)d_tot; A.nonunique = nonunique;
    TB_CUDA(ctx, tb::launch_anchor_count(X, A, (int)wn, st));
    ctx->launches++;
    std::vector<unsigned long long> tot(2 * wn);
    TB_CUDA(ctx, cudaMemcpyAsync(tot.data(), d_tot, wn * 16, cudaMemcpyDeviceToHost, st));
    TB_CUDA(ctx, cudaStreamSynchronize(st));
    std::vector<long long> off; std::vector<unsigned> size;
    size_t elems = 0, take = 0;
    for (; take < wn; ++take) {
      const unsigned s0 = next_pow2(2 * tot[2 * take]), s1 = next_pow2(2 * tot[2 * take + 1]);
      if (take > 0 && elems + s0 + s1 > budget_elems) break;
      off.push_back((long long)elems); size.push_back(s0); elems += s0;
      off.push_back((long long)elems); size.push_back(s1); elems += s1;
    }
    void *d_off, *d_size, *d_keys, *d_cnt;
    TB_CUDA(ctx, S.up(&d_off, off.data(), off.size() * 8)); TB_CUDA(ctx, S.up(&d_size, size.data(), size.size() * 4));
    TB_CUDA(ctx, S.alloc(&d_keys, elems * 8)); TB_CUDA(ctx, S.alloc(&d_cnt, elems * 4));
    A.tab_keys = (long long*)d_keys; A.tab_cnt = (unsigned*)d_cnt; A.tab_off = (const long long*)d_off; A.tab_size = (const unsigned*)d_size;
    TB_CUDA(ctx, tb::launch_anchor_fill(X, A, (int)take, (long long)elems, st));
    ctx->launches += 3;
    TB_CUDA(ctx, cudaStreamSynchronize(st));
    done += take;
  }
  return TB_OK;
}
}  // namespace

int tb_anchor(tb_ctx* ctx, const tb_index* idx, const tb_arena* cons, size_t ntraces, int32_t mem, tb_anchor_config cfg, tb_anchor_result* res) {
  if (!ctx) return TB_ERR_INVALID;
  if (!idx || !cons || !res) return fail(ctx, TB_ERR_INVALID, "null index/arena/result");
  if (ntraces == 0) return TB_OK;
  if (ntraces > (size_t)INT_MAX) return fail(ctx, TB_ERR_INVALID, "ntraces too large");
  if (!cons->base || !cons->off || !cons->len || !res->anchored || !res->forward || !res->kmersupport || !res->bestpos)
    return fail(ctx, TB_ERR_INVALID, "null pointer in anchoring call");
  if (cfg.kmer < 1 || cfg.kmer > 16) return fail(ctx, TB_ERR_UNSUPPORTED, "kmer must be 1..16 (index depth is 16 characters)");
  if (cfg.trim_left < 0 || cfg.trim_right < 0 || cfg.trim_left > 65535 || cfg.trim_right > 65535 || cfg.min_kmer_support < 0)
    return fail(ctx, TB_ERR_INVALID, "trim/support out of range");
  if (mem != TB_MEM_HOST && mem != TB_MEM_DEVICE) return fail(ctx, TB_ERR_INVALID, "mem must be TB_MEM_HOST or TB_MEM_DEVICE");
  if (idx->device != ctx->device) return fail(ctx, TB_ERR_INVALID, "index lives on another device");
  TB_CUDA(ctx, cudaSetDevice(ctx->device));
  Lane& L = ctx->lanes[0];
  cudaStream_t st = L.stream;
  const size_t nt = ntraces;
  std::vector<int32_t> hlen_dev;
  const int32_t* hlen = cons->len;
  if (mem == TB_MEM_DEVICE) {
    hlen_dev.resize(nt);
    TB_CUDA(ctx, cudaMemcpy(hlen_dev.data(), cons->len, nt * 4, cudaMemcpyDeviceToHost));
    hlen = hlen_dev.data();
  }
  int maxlen = 0;
  for (size_t i = 0; i < nt; ++i) {
    if (hlen[i] < 0 || hlen[i] > 65535) return fail(ctx, TB_ERR_UNSUPPORTED, "consensus length outside 0..65535 (the reference scans with a uint16 index)");
    maxlen = std::max(maxlen, hlen[i]);
  }
  tb::KmerIndexView X{idx->rec, idx->n, idx->dir, idx->dir_chars};
  tb::AnchorBatch A{};
  A.trim_left = cfg.trim_left; A.trim_right = cfg.trim_right; A.kmer = cfg.kmer; A.min_support = cfg.min_kmer_support;
  Staged S(st);
  void *d_pass = nullptr;
  uint8_t *o_anch = res->anchored, *o_fwd = res->forward; uint32_t* o_sup = res->kmersupport; int64_t* o_pos = res->bestpos;
  if (mem == TB_MEM_DEVICE) {
    A.cons_base = (const char*)cons->base; A.cons_off = cons->off; A.cons_len = cons->len;
    A.anchored = o_anch; A.forward = o_fwd; A.kmersupport = o_sup; A.bestpos = o_pos;
    if (res->pass) d_pass = res->pass; else TB_CUDA(ctx, S.alloc(&d_pass, nt));
  } else {
    long long cmax = 0;
    for (size_t i = 0; i < nt; ++i) {
      if (cons->off[i] < 0) return fail(ctx, TB_ERR_INVALID, "negative arena offset");
      cmax = std::max<long long>(cmax, cons->off[i] + hlen[i]);
    }
    void *d_c, *d_off, *d_len, *d_a, *d_f, *d_s, *d_p;
    TB_CUDA(ctx, S.up(&d_c, cons->base, (size_t)std::max<long long>(cmax, 1))); TB_CUDA(ctx, S.up(&d_off, cons->off, nt * 8)); TB_CUDA(ctx, S.up(&d_len, cons->len, nt * 4));
    TB_CUDA(ctx, S.alloc(&d_a, nt)); TB_CUDA(ctx, S.alloc(&d_f, nt)); TB_CUDA(ctx, S.alloc(&d_s, nt * 4)); TB_CUDA(ctx, S.alloc(&d_p, nt * 8));
    TB_CUDA(ctx, S.alloc(&d_pass, nt));
    ctx->h2d += (size_t)cmax + nt * 12;
    A.cons_base = (const char*)d_c; A.cons_off = (const int64_t*)d_off; A.cons_len = (const int32_t*)d_len;
    A.anchored = (uint8_t*)d_a; A.forward = (uint8_t*)d_f; A.kmersupport = (uint32_t*)d_s; A.bestpos = (int64_t*)d_p;
  }
  A.pass = (uint8_t*)d_pass;
  // shared-memory tables sized for the longest scan of the batch: two slots per k-mer, 1024..8192 slots (12 B each)
  const unsigned tsize = std::min(std::max(next_pow2(2ull * (unsigned)maxlen), 1024u), 8192u);
  TB_CUDA(ctx, cudaEventRecord(L.k0, st));
  TB_CUDA(ctx, tb::launch_anchor_unique(X, A, (int)nt, tsize, st));
  ctx->launches++;
  TB_CUDA(ctx, cudaEventRecord(L.k1, st));
  std::vector<uint8_t> pass(nt);
  TB_CUDA(ctx, cudaMemcpyAsync(pass.data(), d_pass, nt, cudaMemcpyDeviceToHost, st));
  TB_CUDA(ctx, cudaStreamSynchronize(st));
  TB_CUDA(ctx, cudaEventElapsedTime(&ctx->last_anchor_ms, L.k0, L.k1));
  std::vector<int32_t> todo;
  for (size_t i = 0; i < nt; ++i) if (pass[i] == 3) todo.push_back((int32_t)i);
  if (!todo.empty()) {                                    // scans too long for shared-memory tables: unique pass on global tables
    if (int rc = anchor_global_pass(ctx, X, A, todo, false, st)) return rc;
    TB_CUDA(ctx, cudaMemcpyAsync(pass.data(), d_pass, nt, cudaMemcpyDeviceToHost, st));
    TB_CUDA(ctx, cudaStreamSynchronize(st));
  }
  todo.clear();
  for (size_t i = 0; i < nt; ++i) if (pass[i] == 2) todo.push_back((int32_t)i);
  if (!todo.empty())                                      // "Try using non-unique matches", src/fmindex.h:262-268
    if (int rc = anchor_global_pass(ctx, X, A, todo, true, st)) return rc;
  if (mem == TB_MEM_HOST) {
    TB_CUDA(ctx, cudaMemcpyAsync(o_anch, A.anchored, nt, cudaMemcpyDeviceToHost, st));
    TB_CUDA(ctx, cudaMemcpyAsync(o_fwd, A.forward, nt, cudaMemcpyDeviceToHost, st));
    TB_CUDA(ctx, cudaMemcpyAsync(o_sup, A.kmersupport, nt * 4, cudaMemcpyDeviceToHost, st));
    TB_CUDA(ctx, cudaMemcpyAsync(o_pos, A.bestpos, nt * 8, cudaMemcpyDeviceToHost, st));
    if (res->pass) TB_CUDA(ctx, cudaMemcpyAsync(res->pass, d_pass, nt, cudaMemcpyDeviceToHost, st));
    ctx->d2h += nt * 15;
    TB_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return TB_OK;
}

int tb_ctx_last_anchor_ms(const tb_ctx* ctx, float* unique_ms) {
  if (!ctx || !unique_ms) return TB_ERR_INVALID;
  *unique_ms = ctx->last_anchor_ms;
  return TB_OK;
}

// getReferenceSlice's slice arithmetic, reference src/fmindex.h:286-299 (host).
int tb_reference_slice(int64_t bestpos, const uint32_t* seqlen, int32_t nseq, int32_t conslen, int32_t maxindel,
                       int32_t* refindex, uint32_t* chrpos_out, uint32_t* slicestart, uint32_t* sliceend) {
  if (!seqlen || nseq <= 0 || conslen < 0 || maxindel < 0 || !refindex || !slicestart || !sliceend) return TB_ERR_INVALID;
  int64_t cumsum = 0;
  int32_t ri = 0;
  for (; bestpos >= cumsum + (int64_t)seqlen[ri]; ++ri) {
    cumsum += seqlen[ri];
    if (ri + 1 >= nseq) return TB_ERR_INVALID;            // the reference would run off seqlen here
  }
  const int64_t chrpos_signed = bestpos - cumsum;
  const uint32_t chrpos = chrpos_signed > 0 ? (uint32_t)chrpos_signed : 0;
  uint32_t s0 = 0, s1 = seqlen[ri];
  if (chrpos > (uint32_t)maxindel) s0 = chrpos - (uint32_t)maxindel;
  const uint32_t tmpend = chrpos + (uint32_t)conslen + (uint32_t)maxindel;
  if (tmpend < seqlen[ri]) s1 = tmpend;
  *refindex = ri; *slicestart = s0; *sliceend = s1;
  if (chrpos_out) *chrpos_out = chrpos;
  return TB_OK;
}

// ---- allelicFraction (kernel in fraction.cu) ---------------------------------------------
int tb_allelic_fraction(tb_ctx* ctx, const tb_fraction_batch* b, double* a1, double* a2) {
  if (!ctx) return TB_ERR_INVALID;
  if (!b || !a1 || !a2) return fail(ctx, TB_ERR_INVALID, "null batch/output");
  const size_t nt = b->ntraces;
  if (nt == 0) return TB_OK;
  if (nt > (size_t)INT_MAX) return fail(ctx, TB_ERR_INVALID, "ntraces too large");
  const bool tset = (b->mem & TB_TRACE_SET) != 0;
  if (!b->trace.base || (!tset && (!b->trace.off || !b->trace.len)) || !b->bcpos.base || !b->bcpos.off || !b->bcpos.len || !b->primary_base || !b->secdecompose_base)
    return fail(ctx, TB_ERR_INVALID, "null pointer in fraction batch");
  if (b->trim_left < 0 || b->trim_right < 0) return fail(ctx, TB_ERR_INVALID, "negative trim");
  if ((b->mem & ~TB_TRACE_SET) != TB_MEM_HOST && b->mem != TB_MEM_DEVICE)
    return fail(ctx, TB_ERR_INVALID, "batch->mem must be TB_MEM_HOST (optionally with TB_TRACE_SET) or TB_MEM_DEVICE");
  TB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->lanes[0].stream;
  // the grid values exactly as `for (double i = 0; i <= 1; i += 0.01)` produces them (src/decompose.h:581-585)
  std::vector<double> grid;
  for (double v = 0; v <= 1; v += 0.01) grid.push_back(v);
  if (grid.size() > 101) return fail(ctx, TB_ERR_CUDA, "unexpected grid length");
  std::vector<int32_t> hl;
  const int32_t* bl = b->bcpos.len;
  const int32_t* tl = b->trace.len;
  std::vector<int32_t> hl2;
  Staged S(st);
  TraceView V;
  if (b->mem == TB_MEM_DEVICE) {
    hl.resize(nt); hl2.resize(nt);
    TB_CUDA(ctx, cudaMemcpy(hl.data(), b->bcpos.len, nt * 4, cudaMemcpyDeviceToHost));
    TB_CUDA(ctx, cudaMemcpy(hl2.data(), b->trace.len, nt * 4, cudaMemcpyDeviceToHost));
    bl = hl.data(); tl = hl2.data();
  } else {
    if (int rc = resolve_traces(ctx, b->trace, b->mem, nt, S, V)) return rc;
    tl = V.host_len;
  }
  int maxD = 1;
  for (size_t i = 0; i < nt; ++i) {
    if (bl[i] < 0 || tl[i] < 0 || (bl[i] > 0 && tl[i] == 0)) return fail(ctx, TB_ERR_INVALID, "negative length / empty trace");
    maxD = std::max(maxD, bl[i]);
  }
  if ((size_t)4 * maxD * 9 + 16 > 200 * 1024) return fail(ctx, TB_ERR_UNSUPPORTED, "more than 5600 basecalls in one trace");
  tb::FractionBatch F{};
  F.trim_left = b->trim_left; F.trim_right = b->trim_right; F.ngrid = (int)grid.size();
  void *d_grid, *d_status;
  TB_CUDA(ctx, S.up(&d_grid, grid.data(), grid.size() * 8));
  TB_CUDA(ctx, S.alloc(&d_status, nt));
  F.grid = (const double*)d_grid; F.status = (uint8_t*)d_status;
  if (b->mem == TB_MEM_DEVICE) {
    F.trace_base = (const int32_t*)b->trace.base; F.trace_off = b->trace.off; F.trace_len = b->trace.len;
    F.bcpos_base = (const int32_t*)b->bcpos.base; F.bc_off = b->bcpos.off; F.bc_len = b->bcpos.len;
    F.pri_base = b->primary_base; F.sec_base = b->secdecompose_base; F.a1 = a1; F.a2 = a2;
  } else {
    long long bmax = 0;
    for (size_t i = 0; i < nt; ++i) {
      if (b->bcpos.off[i] < 0) return fail(ctx, TB_ERR_INVALID, "negative offset");
      bmax = std::max<long long>(bmax, b->bcpos.off[i] + bl[i]);
    }
    const long long tmax = (long long)(V.uploaded / 4);
    const void *d_tr = V.base, *d_toff = V.off, *d_tlen = V.len;
    void *d_bp, *d_pri, *d_sec, *d_boff, *d_blen, *d_a1, *d_a2;
    TB_CUDA(ctx, S.up(&d_bp, b->bcpos.base, (size_t)std::max(bmax, 1ll) * 4));
    TB_CUDA(ctx, S.up(&d_pri, b->primary_base, (size_t)std::max(bmax, 1ll))); TB_CUDA(ctx, S.up(&d_sec, b->secdecompose_base, (size_t)std::max(bmax, 1ll)));
    TB_CUDA(ctx, S.up(&d_boff, b->bcpos.off, nt * 8)); TB_CUDA(ctx, S.up(&d_blen, b->bcpos.len, nt * 4));
    TB_CUDA(ctx, S.alloc(&d_a1, nt * 8)); TB_CUDA(ctx, S.alloc(&d_a2, nt * 8));
    ctx->h2d += (size_t)tmax * 4 + (size_t)bmax * 6 + nt * 24;
    F.trace_base = (const int32_t*)d_tr; F.trace_off = (const int64_t*)d_toff; F.trace_len = (const int32_t*)d_tlen;
    F.bcpos_base = (const int32_t*)d_bp; F.bc_off = (const int64_t*)d_boff; F.bc_len = (const int32_t*)d_blen;
    F.pri_base = (const char*)d_pri; F.sec_base = (const char*)d_sec; F.a1 = (double*)d_a1; F.a2 = (double*)d_a2;
  }
  Lane& L = ctx->lanes[0];
  TB_CUDA(ctx, cudaEventRecord(L.k0, st));
  TB_CUDA(ctx, tb::launch_allelic_fraction(F, (int)nt, maxD, st));
  ctx->launches++;
  TB_CUDA(ctx, cudaEventRecord(L.k1, st));
  if (b->mem != TB_MEM_DEVICE) {
    TB_CUDA(ctx, cudaMemcpyAsync(a1, F.a1, nt * 8, cudaMemcpyDeviceToHost, st));
    TB_CUDA(ctx, cudaMemcpyAsync(a2, F.a2, nt * 8, cudaMemcpyDeviceToHost, st));
    ctx->d2h += nt * 16;
  }
  TB_CUDA(ctx, cudaStreamSynchronize(st));
  TB_CUDA(ctx, cudaEventElapsedTime(&ctx->last_fraction_ms, L.k0, L.k1));
  return TB_OK;
}

int tb_ctx_last_fraction_ms(const tb_ctx* ctx, float* ms) {
  if (!ctx || !ms) return TB_ERR_INVALID;
  *ms = ctx->last_fraction_ms;
  return TB_OK;
}

// ---- trace-file ingest (scan + kernel in trace_io.cu) -------------------------------------
int tb_trace_scan(const uint8_t* files, const int64_t* off, const int64_t* len, size_t nfiles, tb_trace_info* info) {
  if (!files || !off || !len || !info) return TB_ERR_INVALID;
  for (size_t i = 0; i < nfiles; ++i) {
    if (off[i] < 0 || len[i] < 0) return TB_ERR_INVALID;
    tb::TraceDesc d;
    tb::scan_trace_file(files + off[i], len[i], &d);
    info[i].format = d.format; info[i].ok = d.ok; info[i].status = d.status; info[i].nsamples = d.ns; info[i].nbasecalls = d.nb;
  }
  return TB_OK;
}

int tb_trace_unpack(tb_ctx* ctx, const uint8_t* files, const int64_t* off, const int64_t* len, size_t nfiles, int32_t out_mem,
                    int32_t* samples, const int64_t* samples_off, int32_t* ploc, uint8_t* qual, char* basecalls1, char* basecalls2,
                    const int64_t* bc_off) {
  if (!ctx) return TB_ERR_INVALID;
  if (!files || !off || !len || !samples || !samples_off || !ploc || !qual || !basecalls1 || !basecalls2 || !bc_off)
    return fail(ctx, TB_ERR_INVALID, "null pointer in trace unpack");
  if (nfiles == 0) return TB_OK;
  if (nfiles > 65535u * 1024u) return fail(ctx, TB_ERR_INVALID, "too many files in one call");
  if (out_mem != TB_MEM_HOST && out_mem != TB_MEM_DEVICE) return fail(ctx, TB_ERR_INVALID, "out_mem must be TB_MEM_HOST or TB_MEM_DEVICE");
  TB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->lanes[0].stream;
  std::vector<tb::TraceDesc> desc(nfiles);
  long long fmin = LLONG_MAX, fmax = 0, smax = 0, bmax = 0;
  for (size_t i = 0; i < nfiles; ++i) {
    if (off[i] < 0 || len[i] < 0 || samples_off[i] < 0 || bc_off[i] < 0) return fail(ctx, TB_ERR_INVALID, "negative offset/length");
    tb::scan_trace_file(files + off[i], len[i], &desc[i]);
    fmin = std::min<long long>(fmin, off[i]); fmax = std::max<long long>(fmax, off[i] + len[i]);
    if (desc[i].format >= 0 && desc[i].status == 0) {
      smax = std::max<long long>(smax, samples_off[i] + 4ll * desc[i].ns);
      bmax = std::max<long long>(bmax, bc_off[i] + desc[i].nb);
    }
  }
  std::vector<int64_t> rel(nfiles);
  for (size_t i = 0; i < nfiles; ++i) rel[i] = off[i] - fmin;
  Staged S(st);
  void *d_files, *d_off, *d_desc, *d_soff, *d_boff;
  TB_CUDA(ctx, S.up(&d_files, files + fmin, (size_t)std::max(fmax - fmin, 1ll)));
  TB_CUDA(ctx, S.up(&d_off, rel.data(), nfiles * 8)); TB_CUDA(ctx, S.up(&d_desc, desc.data(), nfiles * sizeof(tb::TraceDesc)));
  TB_CUDA(ctx, S.up(&d_soff, samples_off, nfiles * 8)); TB_CUDA(ctx, S.up(&d_boff, bc_off, nfiles * 8));
  ctx->h2d += (size_t)(fmax - fmin) + nfiles * (24 + sizeof(tb::TraceDesc));
  tb::TraceUnpack U{};
  U.files = (const uint8_t*)d_files; U.file_off = (const int64_t*)d_off; U.desc = (const tb::TraceDesc*)d_desc;
  U.samples_off = (const int64_t*)d_soff; U.bc_off = (const int64_t*)d_boff;
  void *o_s = samples, *o_p = ploc, *o_q = qual, *o_1 = basecalls1, *o_2 = basecalls2;
  if (out_mem == TB_MEM_HOST) {
    TB_CUDA(ctx, S.alloc(&o_s, (size_t)std::max(smax, 1ll) * 4)); TB_CUDA(ctx, S.alloc(&o_p, (size_t)std::max(bmax, 1ll) * 4));
    TB_CUDA(ctx, S.alloc(&o_q, (size_t)std::max(bmax, 1ll))); TB_CUDA(ctx, S.alloc(&o_1, (size_t)std::max(bmax, 1ll))); TB_CUDA(ctx, S.alloc(&o_2, (size_t)std::max(bmax, 1ll)));
  }
  U.samples = (int32_t*)o_s; U.ploc = (int32_t*)o_p; U.qual = (uint8_t*)o_q; U.basecalls1 = (char*)o_1; U.basecalls2 = (char*)o_2;
  for (size_t f0 = 0; f0 < nfiles; f0 += 65535) {           // gridDim.x limit is far above this; keep launches moderate
    tb::TraceUnpack V = U;
    const size_t fn = std::min<size_t>(65535, nfiles - f0);
    V.file_off += f0; V.desc += f0; V.samples_off += f0; V.bc_off += f0;
    TB_CUDA(ctx, tb::launch_trace_unpack(V, (int)fn, st));
    ctx->launches++;
  }
  bool all_live = true;
  for (size_t i = 0; i < nfiles; ++i) all_live = all_live && desc[i].format >= 0 && desc[i].status == 0;
  if (out_mem == TB_MEM_HOST && all_live) {                 // every item was written: five bulk copies of the arenas' used extents
    long long smin = LLONG_MAX, bmin = LLONG_MAX;
    for (size_t i = 0; i < nfiles; ++i) { smin = std::min<long long>(smin, samples_off[i]); bmin = std::min<long long>(bmin, bc_off[i]); }
    // holes between items (if the caller left any) would be overwritten with scratch contents: fall back to per-item copies then
    long long sused = 0, bused = 0;
    for (size_t i = 0; i < nfiles; ++i) { sused += 4ll * desc[i].ns; bused += desc[i].nb; }
    if (sused == smax - smin && bused == bmax - bmin) {
      if (sused) TB_CUDA(ctx, cudaMemcpyAsync(samples + smin, (const int32_t*)o_s + smin, (size_t)sused * 4, cudaMemcpyDeviceToHost, st));
      if (bused) {
        TB_CUDA(ctx, cudaMemcpyAsync(ploc + bmin, (const int32_t*)o_p + bmin, (size_t)bused * 4, cudaMemcpyDeviceToHost, st));
        TB_CUDA(ctx, cudaMemcpyAsync(qual + bmin, (const uint8_t*)o_q + bmin, (size_t)bused, cudaMemcpyDeviceToHost, st));
        TB_CUDA(ctx, cudaMemcpyAsync(basecalls1 + bmin, (const char*)o_1 + bmin, (size_t)bused, cudaMemcpyDeviceToHost, st));
        TB_CUDA(ctx, cudaMemcpyAsync(basecalls2 + bmin, (const char*)o_2 + bmin, (size_t)bused, cudaMemcpyDeviceToHost, st));
      }
      ctx->d2h += (size_t)sused * 4 + (size_t)bused * 7;
    } else {
      all_live = false;
    }
  }
  if (out_mem == TB_MEM_HOST && !all_live) {
    for (size_t i = 0; i < nfiles; ++i) {                   // items may be sparse in the caller's arenas
      const tb::TraceDesc& d = desc[i];
      if (d.format < 0 || d.status != 0) continue;
      if (d.ns) TB_CUDA(ctx, cudaMemcpyAsync(samples + samples_off[i], (const int32_t*)o_s + samples_off[i], (size_t)16 * d.ns, cudaMemcpyDeviceToHost, st));
      if (d.nb) {
        TB_CUDA(ctx, cudaMemcpyAsync(ploc + bc_off[i], (const int32_t*)o_p + bc_off[i], (size_t)4 * d.nb, cudaMemcpyDeviceToHost, st));
        TB_CUDA(ctx, cudaMemcpyAsync(qual + bc_off[i], (const uint8_t*)o_q + bc_off[i], (size_t)d.nb, cudaMemcpyDeviceToHost, st));
        TB_CUDA(ctx, cudaMemcpyAsync(basecalls1 + bc_off[i], (const char*)o_1 + bc_off[i], (size_t)d.nb, cudaMemcpyDeviceToHost, st));
        TB_CUDA(ctx, cudaMemcpyAsync(basecalls2 + bc_off[i], (const char*)o_2 + bc_off[i], (size_t)d.nb, cudaMemcpyDeviceToHost, st));
      }
      ctx->d2h += (size_t)16 * d.ns + (size_t)7 * d.nb;
    }
  }
  TB_CUDA(ctx, cudaStreamSynchronize(st));
  return TB_OK;
}
