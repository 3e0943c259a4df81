// Several GPUs of one node behind the C ABI (BASELINE.json configs[4]): one handle owns a context per device.
// Pairs (and traces) are independent, so a batch is cut into contiguous, cost-balanced ranges, every range runs through its
// device's own chunk pipeline (capi.cu run_gotoh) on its own host thread, and each device writes its slice of the caller's result
// arrays -- the "gather" of the scores is that write, there is no data-path collective. The one real exchange is the reference text:
// it crosses PCIe once (to the first device) and then travels GPU to GPU (cudaMemcpyPeerAsync: NVLink / NVSwitch where peers are
// connected) along a doubling tree; every device then builds its own anchoring index from its copy (17 ms per 64 Mbp, cheaper than
// shipping 17 B/char of index).
// Only the public entry points of tracy_b200.h are used per device; this file adds no kernel.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tracy_b200.h"

struct tb_multi {
  std::vector<int> devices;
  std::vector<tb_ctx*> ctx;
  std::string err;
};

namespace {
int mfail(tb_multi* m, int code, const std::string& msg) {
  if (m) m->err = msg;
  return code;
}
// contiguous ranges of equal total cost: first[d] .. first[d+1]
void split_by_cost(const std::vector<double>& cost, int parts, std::vector<size_t>& first) {
  const size_t n = cost.size();
  first.assign((size_t)parts + 1, n);
  first[0] = 0;
  double total = 0;
  for (double c : cost) total += c;
  double acc = 0;
  size_t i = 0;
  for (int d = 1; d < parts; ++d) {
    const double want = total * d / parts;
    while (i < n && acc + cost[i] * 0.5 < want) acc += cost[i++];
    first[(size_t)d] = i;
  }
}
template <typename F>
int run_on_all(tb_multi* m, F&& body) {
  const size_t nd = m->ctx.size();
  std::vector<int> rc(nd, TB_OK);
  std::vector<std::thread> th;
  for (size_t d = 1; d < nd; ++d) th.emplace_back([&, d]() { rc[d] = body((int)d); });
  rc[0] = body(0);
  for (auto& t : th) t.join();
  for (size_t d = 0; d < nd; ++d)
    if (rc[d] != TB_OK) return mfail(m, rc[d], "device " + std::to_string(m->devices[d]) + ": " + tb_last_error(m->ctx[d]));
  return TB_OK;
}
}  // namespace

extern "C" {

int tb_multi_create(tb_multi** out, const int* devices, int ndev) {
  if (!out) return TB_ERR_INVALID;
  *out = nullptr;
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible <= 0) { cudaGetLastError(); return TB_ERR_CUDA; }
  tb_multi* m = new tb_multi();
  if (ndev <= 0) { ndev = visible; devices = nullptr; }
  for (int i = 0; i < ndev; ++i) m->devices.push_back(devices ? devices[i] : i);
  for (int d : m->devices) {
    tb_ctx* c = nullptr;
    const int rc = (d >= 0 && d < visible) ? tb_ctx_create(&c, d) : TB_ERR_INVALID;
    if (rc != TB_OK) { tb_multi_destroy(m); return rc; }
    m->ctx.push_back(c);
  }
  // peer access for the text broadcast (best effort: without it cudaMemcpyPeerAsync stages through the host)
  for (size_t a = 0; a < m->devices.size(); ++a)
    for (size_t b = 0; b < m->devices.size(); ++b) {
      if (m->devices[a] == m->devices[b]) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, m->devices[a], m->devices[b]) == cudaSuccess && can) {
        cudaSetDevice(m->devices[a]);
        if (cudaDeviceEnablePeerAccess(m->devices[b], 0) != cudaSuccess) cudaGetLastError();   // already enabled is fine
      }
    }
  *out = m;
  return TB_OK;
}

void tb_multi_destroy(tb_multi* m) {
  if (!m) return;
  for (tb_ctx* c : m->ctx) tb_ctx_destroy(c);
  delete m;
}

int tb_multi_size(const tb_multi* m) { return m ? (int)m->ctx.size() : 0; }
tb_ctx* tb_multi_ctx(tb_multi* m, int i) { return (m && i >= 0 && (size_t)i < m->ctx.size()) ? m->ctx[(size_t)i] : nullptr; }
const char* tb_multi_last_error(const tb_multi* m) { return m ? m->err.c_str() : "null handle"; }

int tb_multi_partition(const int32_t* len1, const int32_t* len2, size_t n, int parts, size_t* first) {
  if (!first || parts <= 0 || (n && (!len1 || !len2))) return TB_ERR_INVALID;
  std::vector<double> cost(n);
  for (size_t i = 0; i < n; ++i) cost[i] = ((double)len1[i] + 1) * ((double)len2[i] + 1);
  std::vector<size_t> f;
  split_by_cost(cost, parts, f);
  std::copy(f.begin(), f.end(), first);
  return TB_OK;
}

int tb_multi_gotoh(tb_multi* m, int kind, const tb_batch* batch, tb_score sc, tb_align_config ac, tb_result* res, size_t* first_out) {
  if (!m) return TB_ERR_INVALID;
  if (!batch || !res) return mfail(m, TB_ERR_INVALID, "null batch/result");
  if (kind < 0 || kind > 2) return mfail(m, TB_ERR_INVALID, "kind: 0 profile x profile, 1 string x string, 2 profile x string");
  if ((batch->mem & 0xff) != TB_MEM_HOST) return mfail(m, TB_ERR_INVALID, "a multi-device batch lives in host memory (a device batch belongs to one device: use its context)");
  const size_t n = batch->npairs, nd = m->ctx.size();
  std::vector<size_t> first(nd + 1, 0);
  if (n) {
    if (!batch->a1.len || !batch->a2.len) return mfail(m, TB_ERR_INVALID, "null length array");
    std::vector<size_t> f(nd + 1);
    const int rc = tb_multi_partition(batch->a1.len, batch->a2.len, n, (int)nd, f.data());
    if (rc != TB_OK) return mfail(m, rc, "partition");
    first = f;
  }
  if (first_out) std::copy(first.begin(), first.end(), first_out);
  if (!n) return TB_OK;
  return run_on_all(m, [&](int d) {
    const size_t p0 = first[(size_t)d], cnt = first[(size_t)d + 1] - p0;
    if (!cnt) return (int)TB_OK;
    tb_batch b = *batch;
    b.a1.off += p0; b.a1.len += p0; b.a2.off += p0; b.a2.len += p0; b.npairs = cnt;
    tb_result r = *res;
    r.scores += p0;
    if (r.ops) r.ops += (int64_t)p0 * r.ops_stride;
    if (r.ops_len) r.ops_len += p0;
    if (r.row0) r.row0 += (int64_t)p0 * r.rows_stride;
    if (r.row1) r.row1 += (int64_t)p0 * r.rows_stride;
    tb_ctx* c = m->ctx[(size_t)d];
    return kind == 0 ? tb_gotoh_pp(c, &b, sc, ac, &r) : kind == 1 ? tb_gotoh_ss(c, &b, sc, ac, &r) : tb_gotoh_ps(c, &b, sc, ac, &r);
  });
}

int tb_multi_index_build(tb_multi* m, const char* text, int64_t text_len, tb_index** out) {
  if (!m) return TB_ERR_INVALID;
  if (!text || !out || text_len <= 0) return mfail(m, TB_ERR_INVALID, "null/empty text");
  const size_t nd = m->ctx.size();
  for (size_t d = 0; d < nd; ++d) out[d] = nullptr;
  std::vector<char*> copy(nd, nullptr);
  std::vector<cudaStream_t> st(nd, nullptr);
  auto drop = [&]() {
    for (size_t d = 0; d < nd; ++d) {
      cudaSetDevice(m->devices[d]);
      if (copy[d]) cudaFree(copy[d]);
      if (st[d]) cudaStreamDestroy(st[d]);
    }
  };
  auto bad = [&](cudaError_t e, const char* what) { cudaGetLastError(); drop(); return mfail(m, TB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); };
  cudaError_t e;
  for (size_t d = 0; d < nd; ++d) {
    if ((e = cudaSetDevice(m->devices[d])) != cudaSuccess) return bad(e, "cudaSetDevice");
    if ((e = cudaMalloc((void**)&copy[d], (size_t)text_len)) != cudaSuccess) return bad(e, "cudaMalloc(text)");
    if ((e = cudaStreamCreateWithFlags(&st[d], cudaStreamNonBlocking)) != cudaSuccess) return bad(e, "cudaStreamCreate");
  }
  // PCIe once, then a doubling tree of GPU-to-GPU copies: after round r the first 2^r devices hold the text
  cudaSetDevice(m->devices[0]);
  if ((e = cudaMemcpyAsync(copy[0], text, (size_t)text_len, cudaMemcpyHostToDevice, st[0])) != cudaSuccess) return bad(e, "cudaMemcpyAsync(text)");
  if ((e = cudaStreamSynchronize(st[0])) != cudaSuccess) return bad(e, "cudaStreamSynchronize");
  for (size_t have = 1; have < nd; have *= 2) {
    for (size_t s = 0; s < have && have + s < nd; ++s) {
      const size_t t = have + s;
      cudaSetDevice(m->devices[t]);
      if (m->devices[s] == m->devices[t]) e = cudaMemcpyAsync(copy[t], copy[s], (size_t)text_len, cudaMemcpyDeviceToDevice, st[t]);
      else e = cudaMemcpyPeerAsync(copy[t], m->devices[t], copy[s], m->devices[s], (size_t)text_len, st[t]);
      if (e != cudaSuccess) return bad(e, "cudaMemcpyPeerAsync(text)");
    }
    for (size_t s = 0; s < have && have + s < nd; ++s) {
      cudaSetDevice(m->devices[have + s]);
      if ((e = cudaStreamSynchronize(st[have + s])) != cudaSuccess) return bad(e, "cudaStreamSynchronize(peer copy)");
    }
  }
  const int rc = run_on_all(m, [&](int d) { return tb_index_build(m->ctx[(size_t)d], copy[(size_t)d], text_len, TB_MEM_DEVICE, &out[d]); });
  drop();
  if (rc != TB_OK)
    for (size_t d = 0; d < nd; ++d) if (out[d]) { tb_index_destroy(m->ctx[d], out[d]); out[d] = nullptr; }
  return rc;
}

int tb_multi_anchor(tb_multi* m, tb_index* const* idx, const tb_arena* consensus, size_t ntraces, tb_anchor_config cfg, tb_anchor_result* res) {
  if (!m) return TB_ERR_INVALID;
  if (!idx || !consensus || !res || (ntraces && (!consensus->off || !consensus->len))) return mfail(m, TB_ERR_INVALID, "null argument");
  const size_t nd = m->ctx.size();
  std::vector<double> cost(ntraces);
  for (size_t i = 0; i < ntraces; ++i) cost[i] = (double)consensus->len[i] + 1;
  std::vector<size_t> first;
  split_by_cost(cost, (int)nd, first);
  return run_on_all(m, [&](int d) {
    const size_t p0 = first[(size_t)d], cnt = first[(size_t)d + 1] - p0;
    if (!cnt) return (int)TB_OK;
    tb_arena a = *consensus;
    a.off += p0; a.len += p0;
    tb_anchor_result r = *res;
    r.anchored += p0; r.forward += p0; r.kmersupport += p0; r.bestpos += p0;
    if (r.pass) r.pass += p0;
    return tb_anchor(m->ctx[(size_t)d], idx[d], &a, cnt, TB_MEM_HOST, cfg, &r);
  });
}

}  // extern "C"
