// Host-side packer of the device aero image (np_aero_pack_host / np_aero_create).
//
// The reference evaluates 43 ReLU MLPs per nlplant call (envs/models/F16/hifi_F16_AeroData.py:12-37,149-819).
// 21 of them take ONE input (20 alpha-only nets + eta_el(el)); a ReLU MLP of one input is exactly a continuous
// piecewise-linear function of that input, so those nets are converted -- exactly, in double precision, from the
// shipped fp32 weights -- into breakpoint tables: per segment (a0, y0, slope), y(x) = y0 + slope * (x - a0) with
// the input z-score and the output de-normalisation folded in.  The 22 nets with 2 or 3 inputs keep their weights
// (transposed to input-major, every layer 16-byte aligned for LDS.128).
//
// Image layout (32-bit words; see f16_device.cuh for the device-side readers):
//   [0 .. kHdrWords)            header (ints): offsets / level counts below
//   znorm  [kNumZ][2]           (mean, std) per input-normalisation group
//   onorm  [43][2]              (mean, std) per net output
//   weights                     nets 0..21 except eta_el: per layer bias[out], W^T[in][out], padded to 4 words
//   c0     [16]                 outputs of the 16 (alpha,beta) nets at alpha = beta = 0 (filled on device)
//   bp_a   [2^La - 1]           merged, sorted alpha breakpoints (deg) of the 21 alpha nets, +inf padded
//   segmap [(Ma + 1)][24] u8    per merged segment: segment index of each alpha net (21 used bytes)
//   ent_a  float4[]             per net, per segment: (a0, y0, slope, 0); net k starts at taboff[k]
//   bp_e   [2^Le - 1], ent_e    the same for eta_el (input el, deg)
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/nplane.h"
#include "f16_layout.h"

namespace npl {

struct HostLayer {
  int in, out;
  std::vector<double> W, b;  // W[out][in]
};
using HostNet = std::vector<HostLayer>;

inline HostNet host_net(const float* blob, const np_net_desc& d) {
  HostNet net;
  size_t src = (size_t)d.w_off;
  for (int l = 0; l < d.n_layers; ++l) {
    HostLayer L;
    L.in = d.dims[l];
    L.out = d.dims[l + 1];
    L.W.assign(blob + src, blob + src + (size_t)L.in * L.out);
    L.b.assign(blob + src + (size_t)L.in * L.out, blob + src + (size_t)L.in * L.out + L.out);
    src += (size_t)L.in * L.out + L.out;
    net.push_back(std::move(L));
  }
  return net;
}

// pre-activations of hidden layer `l` at normalised input z (1-input nets)
inline std::vector<double> pre_at(const HostNet& net, double z, int l) {
  std::vector<double> h{z};
  for (int q = 0; q <= l; ++q) {
    const HostLayer& L = net[q];
    std::vector<double> pre(L.out);
    for (int j = 0; j < L.out; ++j) {
      double a = L.b[j];
      for (int i = 0; i < L.in; ++i) a += L.W[(size_t)j * L.in + i] * h[i];
      pre[j] = a;
    }
    if (q == l) return pre;
    for (double& v : pre) v = v > 0 ? v : 0;
    h.swap(pre);
  }
  return h;
}

// value and derivative (w.r.t. z) of the net output at z: forward pass on dual numbers
inline void dual_at(const HostNet& net, double z, double* y, double* dy) {
  std::vector<double> h{z}, dh{1.0};
  for (size_t q = 0; q < net.size(); ++q) {
    const HostLayer& L = net[q];
    std::vector<double> pre(L.out), dpre(L.out);
    for (int j = 0; j < L.out; ++j) {
      double a = L.b[j], da = 0;
      for (int i = 0; i < L.in; ++i) {
        a += L.W[(size_t)j * L.in + i] * h[i];
        da += L.W[(size_t)j * L.in + i] * dh[i];
      }
      pre[j] = a;
      dpre[j] = da;
    }
    if (q + 1 < net.size())
      for (int j = 0; j < L.out; ++j)
        if (!(pre[j] > 0)) pre[j] = 0, dpre[j] = 0;
    h.swap(pre);
    dh.swap(dpre);
  }
  *y = h[0];
  *dy = dh[0];
}

struct PwlTable {
  std::vector<double> bp;          // sorted breakpoints in INPUT units (deg)
  std::vector<double> a0, y0, sl;  // bp.size() + 1 segments
};

// Exact piecewise-linear form of a 1-input ReLU net, x in input units: z = (x - mean) / std, out = y * ostd + omean.
inline PwlTable build_pwl(const HostNet& net, double mean, double std, double omean, double ostd) {
  const double L = 1.0e6;  // breakpoints are searched in |x| <= 1e6 deg; beyond, the outermost pieces extend linearly
  std::vector<double> cur{-L, L};
  for (size_t l = 0; l + 1 < net.size(); ++l) {
    std::vector<double> add;
    for (size_t s = 0; s + 1 < cur.size(); ++s) {
      const double a = cur[s], b = cur[s + 1];
      const std::vector<double> pa = pre_at(net, (a - mean) / std, (int)l), pb = pre_at(net, (b - mean) / std, (int)l);
      for (size_t j = 0; j < pa.size(); ++j)
        if ((pa[j] < 0) != (pb[j] < 0) && pa[j] != pb[j]) add.push_back(a + (b - a) * pa[j] / (pa[j] - pb[j]));
    }
    cur.insert(cur.end(), add.begin(), add.end());
    std::sort(cur.begin(), cur.end());
    cur.erase(std::unique(cur.begin(), cur.end()), cur.end());
  }
  PwlTable t;
  t.bp.assign(cur.begin() + 1, cur.end() - 1);
  if (t.bp.empty()) t.bp.push_back(0.0);  // a net that is globally affine: one dummy breakpoint
  const size_t M = t.bp.size();
  for (size_t j = 0; j <= M; ++j) {
    double xm, a0;
    if (j == 0) xm = t.bp[0] - 1.0, a0 = t.bp[0];
    else if (j == M) xm = t.bp[M - 1] + 1.0, a0 = t.bp[M - 1];
    else xm = 0.5 * (t.bp[j - 1] + t.bp[j]), a0 = t.bp[j - 1];
    a0 = (double)(float)a0;  // the device subtracts the fp32-rounded anchor; the piece extends linearly to it
    double y, dy;
    dual_at(net, (xm - mean) / std, &y, &dy);
    t.a0.push_back(a0);
    t.sl.push_back(dy / std * ostd);
    t.y0.push_back((y + dy * ((a0 - xm) / std)) * ostd + omean);
  }
  return t;
}

inline int levels_for(size_t n_bp) {
  int lv = 1;
  while (((size_t)1 << lv) - 1 < n_bp) ++lv;
  return lv;
}

// Build the device image.  Returns an empty string on success, else the error text.
inline std::string pack_aero_image(const float* blob, size_t n_floats, const np_net_desc* descs, const double* norm,
                                   int n_nets, std::vector<uint32_t>* image) {
  if (n_nets != kNumNets) return "expected 43 nets";
  std::vector<float> head(kWeightOff, 0.0f);
  bool zset[kNumZ] = {};
  std::string err;
  auto set_z = [&](int zid, double mean, double sd) {
    const float m = (float)mean, s = (float)sd;
    if (zset[zid] && (head[kZnormOff + 2 * zid] != m || head[kZnormOff + 2 * zid + 1] != s))
      err = "nets of one normalisation group disagree on (mean, std)";
    head[kZnormOff + 2 * zid] = m;
    head[kZnormOff + 2 * zid + 1] = s;
    zset[zid] = true;
  };
  std::vector<float> weights(kWeightFloats, 0.0f);
  std::vector<PwlTable> tabs(kNumNets);
  for (int k = 0; k < kNumNets; ++k) {
    const np_net_desc& d = descs[k];
    const NetArch a = arch_of(k);
    const ZSel z = zsel_of(k);
    const int nl = a.h3 ? 4 : 3;
    const int dims[5] = {a.nin, a.h1, a.h2, a.h3 ? a.h3 : 1, a.h3 ? 1 : 0};
    if (d.n_in != a.nin || d.n_layers != nl) return "net architecture mismatch (depth)";
    for (int l = 0; l <= nl; ++l)
      if (d.dims[l] != dims[l]) return "net architecture mismatch (width)";
    const int want_sel[3] = {z.a >= 0 ? 0 : 2, a.nin >= 2 ? 1 : -1, a.nin == 3 ? 2 : -1};
    for (int j = 0; j < a.nin; ++j)
      if (d.sel[j] != want_sel[j]) return "net input selection mismatch";
    size_t need = (size_t)d.w_off;
    for (int l = 0; l < nl; ++l) need += (size_t)dims[l] * dims[l + 1] + dims[l + 1];
    if (need > n_floats) return "blob too short";
    const double* nm = norm + 8 * k;
    for (int j = 0; j < a.nin; ++j) {
      const int zid = d.sel[j] == 0 ? z.a : (d.sel[j] == 1 ? z.b : z.e);
      set_z(zid, nm[j], nm[3 + j]);
    }
    if (!err.empty()) return err;
    head[kOnormOff + 2 * k] = (float)nm[6];
    head[kOnormOff + 2 * k + 1] = (float)nm[7];
    if (a.nin == 1) {
      tabs[k] = build_pwl(host_net(blob, d), (double)(float)nm[0], (double)(float)nm[3], (double)(float)nm[6],
                          (double)(float)nm[7]);
      if (tabs[k].bp.size() + 1 > 255) return "a one-input net has more than 255 linear pieces";
    } else {
      // weights: source per layer W[out][in] then b[out]; device per layer b[out] then W^T[in][out], padded to 4
      size_t src = (size_t)d.w_off;
      int dst = mlp_offset(k) - kWeightOff;
      for (int l = 0; l < nl; ++l) {
        const int in = dims[l], o = dims[l + 1];
        const float* W = blob + src;
        const float* b = W + (size_t)in * o;
        for (int j = 0; j < o; ++j) weights[dst + j] = b[j];
        for (int i = 0; i < in; ++i)
          for (int j = 0; j < o; ++j) weights[dst + o + i * o + j] = W[(size_t)j * in + i];
        src += (size_t)in * o + o;
        dst += layer_floats(in, o);
      }
    }
  }
  for (int zid = 0; zid < kNumZ; ++zid)
    if (!zset[zid]) return "normalisation group without nets";

  // merged alpha breakpoints + per-net segment map
  std::vector<double> merged;
  for (int k = kFirstA1; k < kNumNets; ++k) merged.insert(merged.end(), tabs[k].bp.begin(), tabs[k].bp.end());
  std::sort(merged.begin(), merged.end());
  // breakpoints are compared in fp32 on the device: collapse values that round to the same float
  std::vector<float> bp_a;
  for (double v : merged)
    if (bp_a.empty() || (float)v != bp_a.back()) bp_a.push_back((float)v);
  if (levels_for(bp_a.size()) > kLevelsA || levels_for(tabs[kEtaEl].bp.size()) > kLevelsE)
    return "too many breakpoints for the fixed-depth table search";
  const int La = kLevelsA, Le = kLevelsE;
  const size_t Ma = bp_a.size();

  std::vector<uint32_t>& im = *image;
  im.assign(kWeightOff + kWeightFloats + kNumAB2, 0u);
  memcpy(im.data(), head.data(), head.size() * 4);
  memcpy(im.data() + kWeightOff, weights.data(), weights.size() * 4);
  auto hdr = [&](int i) -> int32_t& { return reinterpret_cast<int32_t*>(im.data())[i]; };  // im reallocates as it grows
  auto align4 = [&]() { while (im.size() % 4) im.push_back(0u); };
  auto push_f = [&](float f) { uint32_t u; memcpy(&u, &f, 4); im.push_back(u); };
  const float inf = std::numeric_limits<float>::infinity();

  hdr(kHdrC0) = kWeightOff + kWeightFloats;
  hdr(kHdrLevelsA) = La;
  hdr(kHdrBpA) = (int32_t)im.size();
  for (size_t i = 0; i < ((size_t)1 << La) - 1; ++i) push_f(i < Ma ? bp_a[i] : inf);
  align4();
  hdr(kHdrSegmap) = (int32_t)im.size();
  hdr(kHdrNumSegA) = (int32_t)(Ma + 1);
  {
    std::vector<uint8_t> row(kSegmapRowBytes);
    for (size_t m = 0; m <= Ma; ++m) {
      // a representative point strictly inside merged segment m (in fp32 breakpoint space)
      double x;
      if (m == 0) x = (double)bp_a[0] - 1.0;
      else if (m == Ma) x = (double)bp_a[Ma - 1] + 1.0;
      else x = 0.5 * ((double)bp_a[m - 1] + (double)bp_a[m]);
      std::fill(row.begin(), row.end(), 0);
      for (int k = kFirstA1; k < kNumNets; ++k) {
        const std::vector<double>& b = tabs[k].bp;
        row[k - kFirstA1] = (uint8_t)(std::upper_bound(b.begin(), b.end(), x) - b.begin());
      }
      for (int w = 0; w < kSegmapRowBytes / 4; ++w) {
        uint32_t u;
        memcpy(&u, row.data() + 4 * w, 4);
        im.push_back(u);
      }
    }
  }
  align4();
  hdr(kHdrEntA) = (int32_t)im.size();
  {
    int off = 0;
    for (int k = kFirstA1; k < kNumNets; ++k) {
      hdr(kHdrTabOff + (k - kFirstA1)) = off;
      const PwlTable& t = tabs[k];
      for (size_t j = 0; j < t.a0.size(); ++j) {
        push_f((float)t.a0[j]); push_f((float)t.y0[j]); push_f((float)t.sl[j]); push_f(0.0f);
      }
      off += (int)t.a0.size();
    }
  }
  {
    const PwlTable& t = tabs[kEtaEl];
    hdr(kHdrLevelsE) = Le;
    hdr(kHdrBpE) = (int32_t)im.size();
    for (size_t i = 0; i < ((size_t)1 << Le) - 1; ++i) push_f(i < t.bp.size() ? (float)t.bp[i] : inf);
    align4();
    hdr(kHdrEntE) = (int32_t)im.size();
    for (size_t j = 0; j < t.a0.size(); ++j) {
      push_f((float)t.a0[j]); push_f((float)t.y0[j]); push_f((float)t.sl[j]); push_f(0.0f);
    }
  }
  align4();
  hdr(kHdrWordsTotal) = (int32_t)im.size();
  if (im.size() * 4 > (size_t)kMaxAeroBytes) return "aero image exceeds the shared-memory budget";
  return "";
}

}  // namespace npl
