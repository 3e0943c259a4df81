// Trace-file ingest: traceFormat / readab / readscf, reference src/scf.h:19-35,
// src/abif.h:286-405, src/scf.h:38-102, for a batch of files.
//
// The host only walks each file's directory (a few hundred bytes) into a descriptor; the raw file bytes go to the GPU as
// they are (big-endian int16 samples: half the bytes of the int32 arrays the reference builds) and one block per
// (file, channel) converts them: byte swap + widening for ABIF, and for SCF 3.x the two rounds of running sums the
// reference applies with an int16 carry (a mod-2^16 prefix scan). The int32 [4][nsamples] layout it writes is the one
// tb_basecall / tb_create_profile read, so file bytes -> samples -> basecalls -> profile -> DP stays on the device.
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tracy_b200.h"
#include "common.cuh"

namespace tb {

namespace {
inline int32_t be32(const uint8_t* p) { return (int32_t)(((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]); }
inline int16_t be16(const uint8_t* p) { return (int16_t)(((uint16_t)p[0] << 8) | (uint16_t)p[1]); }
}  // namespace

// Directory walk of one file (host). Fills the descriptor; returns nothing the device needs beyond it.
void scan_trace_file(const uint8_t* buf, int64_t n, TraceDesc* d) {
  std::memset(d, 0, sizeof(*d));
  d->format = -1;
  if (n < 4) return;
  if (std::memcmp(buf, "ABIF", 4) == 0) d->format = 0;          // traceFormat, src/scf.h:28-29
  else if (std::memcmp(buf, ".scf", 4) == 0) d->format = 1;
  if (d->format == 0) {
    if (n < 34) { d->status = TB_TRACE_TRUNCATED; return; }
    const int16_t esize = be16(buf + 16);                         // src/abif.h:310-312
    const int32_t nelements = be32(buf + 18), offset = be32(buf + 26);
    if (esize < 28 || nelements < 0 || offset < 0 || (int64_t)offset + (int64_t)nelements * esize > n) { d->status = TB_TRACE_TRUNCATED; return; }
    int32_t data_off[4] = {0, 0, 0, 0}, data_n[4] = {0, 0, 0, 0};
    bool seen_data[4] = {false, false, false, false}, seen_ploc = false, seen_pcon = false;
    std::string order;
    int64_t b1 = 0, b2 = 0;                                       // string lengths (incl. the reference's one extra byte)
    for (int32_t i = 0; i < nelements; ++i) {
      const uint8_t* e = buf + (int64_t)i * esize + offset;
      const std::string name((const char*)e, 4);
      const int32_t number = be32(e + 4);
      int16_t etype = be16(e + 8);
      const int16_t es = be16(e + 10);
      const int32_t ne = be32(e + 12), dsize = be32(e + 16), doffset = be32(e + 20);
      if (name == "PCON") etype = 1;                              // src/abif.h:330
      int64_t ofsraw = (int64_t)i * esize + offset + 20;          // data inline in the directory entry ...
      if (dsize > 4) ofsraw = doffset;                            // ... or at doffset, src/abif.h:339-340
      int64_t total = ofsraw + (int64_t)ne * es + 1;              // the reference takes one byte more than the data
      if (total > n) total = n;
      if (ofsraw < 0 || ofsraw > n || ne < 0 || es < 0) { d->status = TB_TRACE_TRUNCATED; return; }
      const int64_t elen = total - ofsraw;
      if (etype == 2) {
        if (name == "PBAS" && number == 2) { d->b1_off = ofsraw; b1 = elen; }
        else if (name == "P2BA" && number == 1) { d->b2_off = ofsraw; b2 = elen; }
        else if (name == "FWO_" && number == 1) order.assign((const char*)buf + ofsraw, (size_t)elen);
      } else if (etype == 4) {
        if (name == "PLOC" && number == 2) {
          if (seen_ploc || ofsraw + 2ll * ne > n) { d->status = seen_ploc ? TB_TRACE_DUPLICATE : TB_TRACE_TRUNCATED; return; }
          seen_ploc = true; d->ploc_off = ofsraw; d->ploc_n = ne;
        } else if (name == "DATA" && number >= 9 && number <= 12) {
          const int k = number - 9;
          if (seen_data[k] || ofsraw + 2ll * ne > n) { d->status = seen_data[k] ? TB_TRACE_DUPLICATE : TB_TRACE_TRUNCATED; return; }
          seen_data[k] = true; data_off[k] = (int32_t)ofsraw; data_n[k] = ne;
        }
      } else if (etype == 1) {
        if (name == "PCON" && number == 2) {
          if (seen_pcon || ofsraw + ne > n) { d->status = seen_pcon ? TB_TRACE_DUPLICATE : TB_TRACE_TRUNCATED; return; }
          seen_pcon = true; d->q_off = ofsraw; d->q_n = ne;
        }
      }
    }
    // src/abif.h:379-388: every vector is cut to the shortest of them (basecalls2 only counts when present)
    int64_t m1 = b1;
    if (b2) m1 = std::min(b1, b2);
    const int64_t m = std::min<int64_t>(m1, std::min<int64_t>(d->q_n, d->ploc_n));
    d->nb = (int32_t)m;
    d->b1_n = (int32_t)std::min(b1, m); d->b2_n = (int32_t)std::min(b2, m);
    // src/abif.h:391-397: channel i of the file goes to the base FWO_ names at i
    int32_t ns = -1;
    bool ragged = false;
    for (int k = 0; k < 4; ++k) { d->ch_off[k] = 0; d->ch_n[k] = 0; }
    for (size_t i = 0; i < order.size(); ++i) {
      const int k = order[i] == 'A' ? 0 : order[i] == 'C' ? 1 : order[i] == 'G' ? 2 : order[i] == 'T' ? 3 : -1;
      if (k < 0) continue;
      if (i >= 4) { d->status = TB_TRACE_TRUNCATED; return; }     // the reference would index past its four channels
      d->ch_off[k] = data_off[i]; d->ch_n[k] = data_n[i];
    }
    for (int k = 0; k < 4; ++k) { if (ns < 0) ns = d->ch_n[k]; else if (d->ch_n[k] != ns) ragged = true; }
    if (ragged) { d->status = TB_TRACE_RAGGED; return; }
    d->ns = ns;
    d->ok = m > 0;                                                // "File lacks basecalls!" otherwise
  } else if (d->format == 1) {
    if (n < 40) { d->status = TB_TRACE_TRUNCATED; return; }
    const int32_t num = be32(buf + 4), offset = be32(buf + 8), nbases = be32(buf + 12), bases_off = be32(buf + 24);
    // version: lexical_cast<float>("3.00") > 2.9, src/scf.h:61-62 (digits '.' digits)
    char v[5] = {(char)buf[36], (char)buf[37], (char)buf[38], (char)buf[39], 0};
    const float nv = strtof(v, nullptr);
    if (num < 0 || nbases < 0 || offset < 0 || bases_off < 0 || (int64_t)offset + 8ll * num > n) { d->status = TB_TRACE_TRUNCATED; return; }
    d->ns = num;
    if (nv > 2.9) {
      if ((int64_t)bases_off + 4ll * nbases > n) { d->status = TB_TRACE_TRUNCATED; return; }
      d->scf_v3 = 1;
      for (int k = 0; k < 4; ++k) { d->ch_off[k] = offset + 2 * k * num; d->ch_n[k] = num; }
      d->ploc_off = bases_off; d->ploc_n = nbases; d->nb = nbases;
      d->ok = 1;                                                  // basecalls stay empty, qualities are zeros (src/scf.h:85-90)
    } else {
      d->scf_v3 = 0;                                              // "SCF version greater 2.9 required!" -> false
      for (int k = 0; k < 4; ++k) { d->ch_off[k] = offset + 2 * k; d->ch_n[k] = num; }
      d->ok = 0;
    }
  }
}

// One block per (file, lane): lanes 0..3 = channels A,C,G,T, lane 4 = basecall positions, qualities, basecall strings.
__global__ void __launch_bounds__(256) trace_unpack_kernel(const TraceUnpack U) {
  const int f = blockIdx.x, lane = blockIdx.y;
  const TraceDesc d = U.desc[f];
  if (d.format < 0 || d.status != 0) return;
  const uint8_t* buf = U.files + U.file_off[f];
  if (lane < 4) {
    int32_t* out = U.samples + U.samples_off[f] + (size_t)lane * d.ns;
    const uint8_t* src = buf + d.ch_off[lane];
    if (d.format == 0 || !d.scf_v3) {
      const int stride = d.format == 0 ? 2 : 8;                   // SCF < 3: samples interleaved A,C,G,T (src/scf.h:80)
      for (int p = threadIdx.x; p < d.ch_n[lane]; p += blockDim.x)
        out[p] = (int16_t)(((unsigned)src[(size_t)p * stride] << 8) | src[(size_t)p * stride + 1]);
      return;
    }
    // SCF 3.x, src/scf.h:66-77: twice  t[p] += prev; prev = (int16) t[p].  The carry lives mod 2^16, so
    // t'[p] = t[p] + sext16(sum_{q<p} t[q] mod 2^16): an exclusive prefix sum per round.
    __shared__ unsigned warp_sum[8];
    __shared__ unsigned carry;
    const int n = d.ch_n[lane];
    for (int p = threadIdx.x; p < n; p += blockDim.x)
      out[p] = (int16_t)(((unsigned)src[(size_t)p * 2] << 8) | src[(size_t)p * 2 + 1]);
    __syncthreads();
    for (int round = 0; round < 2; ++round) {
      if (threadIdx.x == 0) carry = 0;
      __syncthreads();
      for (int base = 0; base < n; base += blockDim.x) {
        const int p = base + threadIdx.x;
        const unsigned v = p < n ? (unsigned)out[p] : 0u;
        unsigned incl = v;
        for (int s = 1; s < 32; s <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, incl, s); if ((threadIdx.x & 31) >= s) incl += o; }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned before = carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += warp_sum[w];
        const unsigned excl = before + incl - v;                  // sum of everything before p
        if (p < n) out[p] = (int32_t)v + (int32_t)(int16_t)(excl & 0xffffu);
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + incl;
        __syncthreads();
      }
    }
    return;
  }
  const int64_t bo = U.bc_off[f];
  for (int p = threadIdx.x; p < d.nb; p += blockDim.x) {
    const uint8_t* s = buf + d.ploc_off;
    U.ploc[bo + p] = d.format == 0 ? (int32_t)(int16_t)(((unsigned)s[2 * p] << 8) | s[2 * p + 1])
                                   : (int32_t)(((unsigned)s[4 * p] << 24) | ((unsigned)s[4 * p + 1] << 16) | ((unsigned)s[4 * p + 2] << 8) | s[4 * p + 3]);
    U.qual[bo + p] = d.format == 0 ? buf[d.q_off + p] : (uint8_t)0;
    // replaceNonDna (src/abif.h:276-284); strings shorter than nb are padded with '\0' by std::string::resize
    char c1 = 0, c2 = 0;
    if (p < d.b1_n) { c1 = (char)buf[d.b1_off + p]; if (c1 != 'A' && c1 != 'C' && c1 != 'G' && c1 != 'T') c1 = 'N'; }
    if (p < d.b2_n) { c2 = (char)buf[d.b2_off + p]; if (c2 != 'A' && c2 != 'C' && c2 != 'G' && c2 != 'T') c2 = 'N'; }
    U.basecalls1[bo + p] = c1;
    U.basecalls2[bo + p] = c2;
  }
}

cudaError_t launch_trace_unpack(const TraceUnpack& U, int nfiles, cudaStream_t st) {
  trace_unpack_kernel<<<dim3(nfiles, 5), 256, 0, st>>>(U);
  return cudaGetLastError();
}

}  // namespace tb
