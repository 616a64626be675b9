// TEST INFRASTRUCTURE -- not part of the product.
//
// Host C++/OpenMP restatement ("port") of the reference's stixel path, one
// image column per loop iteration, for (a) the no-GPU unit tests, (b) the
// cpu_baseline leg of bench.py.  It follows, function by function:
//   Stixels::Initialize/Precompute*      InstanceStixels/src/Stixels.cu:43-248, 786-887
//   JoinColumns                          InstanceStixels/src/StixelsKernels.cu:980-1095
//   ComputeObjectLUT (+warp scan order)  StixelsKernels.cu:236-296, 959-978
//   ComputePrefixSum (Blelloch order)    include/InstanceStixels/StixelsKernels.h:73-103
//   StixelsKernel<PAIRWISE>              StixelsKernels.cu:298-957  (cost helpers :31-234)
//   semantic helpers                     include/InstanceStixels/Cityscapes.h:28-123
//   ClusterInstances/GetInstanceStixels  Stixels.cu:639-681, 744-776 (dbscan_def.h)
//
// Parity status: PINNED against outputs of the reference itself, produced by
// oracle/_ref (the unmodified reference CUDA sources built for sm_100a) on the
// GPU box and committed as tests/golden/*.npz with the generating script
// tools/gpu_parity_report.py.  The reference has no usable golden vectors of
// its own (its Catch2 tests are disabled, CMakeLists.txt:267-280).  It is NOT
// bit-exact with the GPU: the reference's fast-math build uses MUFU.RCP /
// MUFU.LG2 approximations that have no host equivalent (1/x and log2f are used
// here), so comparisons against GPU results use a tolerance; summation orders
// (Blelloch tree, chunked Kogge-Stone), FMA contraction shapes taken from the
// reference's sm_100a SASS and flush-to-zero are reproduced.
// The DBSCAN step is UNPINNED w.r.t. cuML (see dbscan_def.h).
#include <immintrin.h>
#include <omp.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/instance_stixels_b200.h"
#include "dbscan_def.h"

namespace {

constexpr float kInf = std::numeric_limits<float>::infinity();
constexpr float kPi = 3.1416f;        // Stixels.hpp:37
constexpr int kLogLut = 1000000;      // configuration.h:30
constexpr int kMaxSections = 200;
constexpr float kLn2 = 0.69314718246459960938f;
constexpr float kNegLog07 = 0.35667496919631958008f;
constexpr float kLg2_03 = -1.7369655370712280273f;
constexpr int GROUND = 0, OBJECT = 1, SKY = 2;

inline float ffma(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
inline float rcp(float x) { return 1.0f / x; }          // GPU: MUFU.RCP
inline float lg2(float x) { return log2f(x); }          // GPU: MUFU.LG2
inline float neg_log_div(float a, float b) {            // StixelsKernels.cu:35-38
  const float t = lg2(a) * kLn2;
  return ffma(lg2(b), kLn2, -t);
}
inline float clamp_neg(float x) { return x < 0.0f ? 0.0f : x; }

struct Model {
  isx_config cfg;
  int rows, cols, realcols, max_dis, step, margin, hs2, rows_p2;
  float invalid, max_disf, max_dis_log, rows_log, puniform, puniform_sky;
  float pn_sky_log, nopn_sky_log, pn_ground_log, nopn_ground_log, pn_object_log, nopn_object_log;
  float norm_sky, inv_s2_sky, iw, pw, dw, sw, sigma_tilt;
  std::vector<float> log_lut, obj_cost_lut, odr;

  float fast_log(float v) const {  // Stixels.cu:786-788
    long i = (long)(v * kLogLut + 0.5f);
    i = i < 0 ? 0 : (i > kLogLut ? kLogLut : i);
    return log_lut[(size_t)i];
  }

  explicit Model(const isx_config &c) : cfg(c) {
    rows = (int)c.rows; cols = (int)c.cols; max_dis = c.max_dis; step = c.column_step; margin = c.width_margin;
    invalid = c.invalid_disparity;
    realcols = (cols - margin) / step;
    max_disf = (float)max_dis;
    rows_p2 = (int)powf(2, ceilf(log2f(rows + 1)));
    hs2 = (int)powf(2, ceilf(log2f(rows / 8 + 1)));
    pw = c.prior_weight; dw = c.disparity_weight; sw = c.segmentation_weight;
    iw = 0.0;  // Stixels.cu:408-423
    if (c.segmentation_weight > 1e-5) {
      iw = c.instance_weight / c.segmentation_weight;
      if (c.instance_weight < 1e-8) iw = 0.0;
    }
    const float pn_ground = (c.pground_given_nexist * c.pnexist_dis) / c.pground;  // :361-373
    const float pn_object = (c.pobject_given_nexist * c.pnexist_dis) / c.pobject;
    const float pn_sky = (c.psky_given_nexist * c.pnexist_dis) / c.psky;
    sigma_tilt = c.sigma_camera_tilt * (kPi) / 180.0f;
    log_lut.resize((size_t)kLogLut + 1);
    for (int i = 0; i < kLogLut; i++) log_lut[i] = logf((float)i / ((float)kLogLut));
    log_lut[kLogLut] = 0.0f;
    max_dis_log = logf(max_disf);
    rows_log = logf((float)rows);
    puniform_sky = max_dis_log - logf(c.pout_sky);
    puniform = max_dis_log - logf(c.pout);
    pn_sky_log = -logf(pn_sky); nopn_sky_log = -logf(1.0f - pn_sky);
    pn_ground_log = -logf(pn_ground); nopn_ground_log = -logf(1.0f - pn_ground);
    pn_object_log = -logf(pn_object); nopn_object_log = -logf(1.0f - pn_object);
    odr.assign(max_dis, 0.0f);  // :879-887
    for (int i = 1; i < max_dis; i++) {
      const float pm = (float)i;
      const float z = (c.baseline * c.focal / pm) + c.range_objects_z;
      odr[i] = pm - (c.baseline * c.focal / z);
    }
    // erff: the reference is host code compiled by nvcc, whose headers add a global erf(float)
    // overload (= erff); plain g++ would pick ::erf(double) here and round differently.
    {  // PrecomputeSky :856-865
      const float s = c.sigma_sky;
      const float a = 0.5f * (erff(max_disf / (s * sqrtf(2.0f))) - erff(0.0f));
      norm_sky = fast_log(a) - logf((1.0f - c.pout_sky) / (s * sqrtf(2.0f * kPi)));
      inv_s2_sky = 1.0f / (2.0f * s * s);
    }
    std::vector<float> norm_o(max_dis), inv_o(max_dis);  // PrecomputeObject :819-840
    for (int d = 0; d < max_dis; d++) {
      const float fn = (float)d;
      const float so = fn * fn * c.range_objects_z / (c.focal * c.baseline);
      const float s = sqrtf(c.sigma_disparity_object * c.sigma_disparity_object + so * so);
      const float a = 0.5f * (erff((max_disf - fn) / (s * sqrtf(2.0f))) - erff((-fn) / (s * sqrtf(2.0f))));
      norm_o[d] = fast_log(a) - fast_log((1.0f - c.pout) / (s * sqrtf(2.0f * kPi)));
      inv_o[d] = 1.0f / (2.0f * s * s);
    }
    obj_cost_lut.assign((size_t)max_dis * max_dis, 0.0f);  // GetDataCostObject :842-854
    for (int fn = 0; fn < max_dis; fn++)
      for (int d = 0; d < max_dis; d++) {
        float cost = pn_object_log;
        if (d != (int)invalid) {
          const float md = (float)(d - fn);
          const float g = norm_o[fn] + md * md * inv_o[fn];
          cost = fminf(puniform, g) + nopn_object_log;
        }
        obj_cost_lut[(size_t)fn * max_dis + d] = cost;
      }
  }

  // PrecomputeGround :790-817; returns flipped vhor
  int ground(const isx_road &r, std::vector<float> &gf, std::vector<float> &norm, std::vector<float> &inv) const {
    const int vhor = rows - r.vhor - 1;
    gf.resize(rows); norm.resize(rows); inv.resize(rows);
    const float fb = (cfg.focal * cfg.baseline) / r.camera_height;
    for (int v = 0; v < rows; v++) {
      const float fn = r.alpha_ground * (float)(vhor - v);
      gf[v] = fn;
      const float x = r.camera_tilt + (float)(vhor - v) / cfg.focal;
      const float s2r = fb * fb *
                        (cfg.sigma_camera_height * cfg.sigma_camera_height * x * x /
                             (r.camera_height * r.camera_height) +
                         sigma_tilt * sigma_tilt);
      const float s = sqrtf(cfg.sigma_disparity_ground * cfg.sigma_disparity_ground + s2r);
      const float a = 0.5f * (erff((max_disf - fn) / (s * sqrtf(2.0f))) - erff((-fn) / (s * sqrtf(2.0f))));
      norm[v] = fast_log(a) - fast_log((1.0f - cfg.pout) / (s * sqrtf(2.0f * kPi)));
      inv[v] = 1.0f / (2.0f * s * s);
    }
    return vhor;
  }
};

// In-place Blelloch exclusive scan, literally StixelsKernels.h:73-103 with the
// thread loop serialised (n = power of two).
template <typename T>
void blelloch(std::vector<T> &arr, int n) {
  int offset = 1;
  for (int d = n >> 1; d > 0; d >>= 1) {
    for (int t = 0; t < d; t++) {
      const int ai = offset * (2 * t + 1) - 1, bi = offset * (2 * t + 2) - 1;
      arr[bi] += arr[ai];
    }
    offset *= 2;
  }
  arr[n - 1] = 0;
  for (int d = 1; d < n; d *= 2) {
    offset >>= 1;
    for (int t = 0; t < d; t++) {
      const int ai = offset * (2 * t + 1) - 1, bi = offset * (2 * t + 2) - 1;
      const T tmp = arr[ai];
      arr[ai] = arr[bi];
      arr[bi] += tmp;
    }
  }
}

struct ColumnWork {
  std::vector<float> d, valid_ps, disp_ps, ground_ps, sky_ps, lut, cost;
  std::vector<int64_t> mx, my, mx2, my2;
  std::vector<int32_t> seg;      // [21][hs2] exclusive prefix (offsets squared first)
  std::vector<int> index;
};

inline int dsum(const int32_t *ps, int vB, int vT) {  // DownsampledSum, Cityscapes.h:28-42
  const int tm = vT % 8, td = vT / 8, bm = vB % 8, bd = vB / 8;
  return (ps[td] - ps[bd]) * 8 + (ps[td + 1] - ps[td]) * (tm + 1) - (ps[bd + 1] - ps[bd]) * bm;
}

struct Ctx {
  const Model &m;
  const std::vector<float> &gf, &norm_g, &inv_g;
  int vhor;
  bool pairwise;
};

inline float mean_of(const ColumnWork &w, const Model &m, int vB, int vT) {  // ComputeMean :47-60
  if (m.invalid >= 0) {
    const float vd = w.valid_ps[vT + 1] - w.valid_ps[vB];
    return vd == 0 ? 0.0f : (w.disp_ps[vT + 1] - w.disp_ps[vB]) * rcp(vd);
  }
  return (w.disp_ps[vT + 1] - w.disp_ps[vB]) * rcp((float)(vT + 1 - vB));
}

inline float inst_cost(const ColumnWork &w, int vB, int vT) {  // :72-86 as compiled (SASS: FMUL, FFMA, FADD, FMUL, FFMA)
  const float sx = (float)(w.mx[vT + 1] - w.mx[vB]), sy = (float)(w.my[vT + 1] - w.my[vB]);
  const float sx2 = (float)(w.mx2[vT + 1] - w.mx2[vB]), sy2 = (float)(w.my2[vT + 1] - w.my2[vB]);
  const float rn = rcp((float)(vT + 1 - vB));
  return ffma(-(sy * sy), rn, sy2 + ffma(-(sx * sx), rn, sx2));
}

void process_column(const Ctx &c, int col, const float *disp_img, const int32_t *seg_in, ColumnWork &w,
                    isx_section *out) {
  const Model &m = c.m;
  const int H = m.rows, D = m.max_dis, hs2 = m.hs2, P2 = m.rows_p2;
  const float invalid = m.invalid;
  // ---- JoinColumns :980-1095 ----
  w.d.assign(H, 0.f);
  for (int row = 0; row < H; row++) {
    const float *px = disp_img + (size_t)row * m.cols + (size_t)col * m.step + m.margin;
    float val;
    if (m.cfg.median_join) {
      float v[16]; int n = 0;
      for (int i = 0; i < m.step; i++) if (!(invalid >= 0 && px[i] == invalid)) v[n++] = px[i];
      if (n == 0) val = invalid;
      else {
        for (int i = 0; i < n / 2 + 1; i++) {
          int mi = i;
          for (int j = i + 1; j < n; j++) if (v[j] < v[mi]) mi = j;
          std::swap(v[i], v[mi]);
        }
        val = v[n / 2];
        if (n % 2 == 0) val = (val + v[n / 2 - 1]) * 0.5f;
      }
    } else if (invalid >= 0) {
      float sum = 0.f; int bad = 0;
      for (int i = 0; i < m.step; i++) { if (px[i] != invalid) sum += px[i]; else bad++; }
      val = bad != m.step ? rcp((float)(m.step - bad)) * sum : invalid;
    } else {
      float sum = 0.f;
      for (int i = 0; i < m.step; i++) sum += px[i];
      val = rcp((float)m.step) * sum;
    }
    w.d[H - 1 - row] = val;
  }
  // ---- ComputeObjectLUT :236-296, 959-978: rows of (P2+1), chunked Kogge-Stone order ----
  const int lstride = P2 + 1;
  w.lut.assign((size_t)D * lstride, 0.f);
  const int n_p2 = (int)powf(2, ceilf(log2f(H)));
  for (int fn = 0; fn < D; fn++) {
    float *arr = &w.lut[(size_t)fn * lstride];
    float add = 0.f;
    arr[0] = 0.f;
    for (int i = 0; i < n_p2; i += 32) {
      float x[32];
      for (int l = 0; l < 32; l++) {
        int dis = 0;
        if (i + l < H) dis = (int)w.d[i + l];
        dis = dis < 0 ? 0 : (dis >= D ? D - 1 : dis);
        x[l] = m.obj_cost_lut[(size_t)fn * D + dis];
      }
      x[0] += add;
      for (int j = 1; j < 32; j *= 2) {
        float nx[32];
        for (int l = 0; l < 32; l++) nx[l] = l >= j ? x[l] + x[l - j] : x[l];
        std::memcpy(x, nx, sizeof x);
      }
      for (int l = 0; l < 32; l++) if (i + l + 1 < lstride) arr[i + l + 1] = x[l];
      add = x[31];
    }
  }
  // ---- load phase :371-446 ----
  w.valid_ps.assign(P2, 0.f); w.disp_ps.assign(P2, 0.f); w.ground_ps.assign(P2, 0.f); w.sky_ps.assign(P2, 0.f);
  w.mx.assign(P2, 0); w.my.assign(P2, 0); w.mx2.assign(P2, 0); w.my2.assign(P2, 0);
  const int K = m.cfg.n_semantic_classes, CH = K + m.cfg.n_offset_channels;
  w.seg.assign(seg_in + (size_t)col * CH * hs2, seg_in + (size_t)(col + 1) * CH * hs2);
  for (int row = 0; row < H; row++) {
    const float d = w.d[row];
    if (invalid >= 0) {
      const int va = d != invalid;
      w.valid_ps[row] = (float)va;
      w.disp_ps[row] = ((float)va) * d;
    } else {
      w.disp_ps[row] = d;
    }
    const int offy = w.seg[(size_t)K * hs2 + row / 8], offx = w.seg[(size_t)(K + 1) * hs2 + row / 8];
    w.mx[row] = (int64_t)((m.step * col + 0.5 * (m.step - 1.0)) + offx + 0.5);
    w.my[row] = (int64_t)(row - offy + 0.5);
    w.mx2[row] = w.mx[row] * w.mx[row];
    w.my2[row] = w.my[row] * w.my[row];
    float sky = 0.f;  // GetDataCostSky :201-215
    if (row >= c.vhor) {
      sky = m.pn_sky_log;
      if (d != invalid) sky = fminf(m.puniform_sky, ffma(d * d, m.inv_s2_sky, m.norm_sky)) + m.nopn_sky_log;
    }
    w.sky_ps[row] = sky;
    float grd = kInf;  // GetDataCostGround :217-234
    if (row < c.vhor) {
      grd = m.pn_ground_log;
      if (d != invalid) {
        const float diff = d - c.gf[row];
        grd = fminf(m.puniform, ffma(diff * diff, c.inv_g[row], c.norm_g[row])) + m.nopn_ground_log;
      }
    }
    w.ground_ps[row] = grd;
  }
  for (int q = 0; q < hs2; q++) {  // squared offsets :411-416
    w.seg[(size_t)K * hs2 + q] *= w.seg[(size_t)K * hs2 + q];
    w.seg[(size_t)(K + 1) * hs2 + q] *= w.seg[(size_t)(K + 1) * hs2 + q];
  }
  // ---- prefix sums :452-469 ----
  if (invalid >= 0) blelloch(w.valid_ps, P2);
  blelloch(w.disp_ps, P2);
  blelloch(w.mx, P2); blelloch(w.my, P2); blelloch(w.mx2, P2); blelloch(w.my2, P2);
  blelloch(w.ground_ps, P2); blelloch(w.sky_ps, P2);
  for (int ch = 0; ch < CH; ch++) {
    std::vector<int32_t> tmp(w.seg.begin() + (size_t)ch * hs2, w.seg.begin() + (size_t)(ch + 1) * hs2);
    blelloch(tmp, hs2);
    std::copy(tmp.begin(), tmp.end(), w.seg.begin() + (size_t)ch * hs2);
  }
  const int32_t *S = w.seg.data();
  auto seg_ground = [&](int vB, int vT) { return fminf((float)dsum(S, vB, vT), (float)dsum(S + hs2, vB, vT)); };
  auto seg_object = [&](int vB, int vT, float ic, float nic, int *cls) {  // Cityscapes.h:61-111
    float best = kInf; int bc = 2;
    for (int k = 2; k < 19; k++) {
      if (k == 10) continue;
      float cs = 0.0f + (k < 10 ? nic : ic);
      cs += (float)dsum(S + (size_t)k * hs2, vB, vT);
      if (best > cs) { best = cs; bc = k; }
    }
    if (cls) *cls = bc;
    return best;
  };
  // ---- DP :477-839 ----
  w.cost.assign((size_t)3 * H, kInf);
  w.index.assign((size_t)3 * H, 0);
  const float pw = m.pw, dw = m.dw, sw = m.sw, iw = m.iw, eps = m.cfg.epsilon;
  const int32_t *ox = S + (size_t)(K + 1) * hs2, *oy = S + (size_t)K * hs2;
  for (int vT = 0; vT < H; vT++) {  // first segment, vB = 0
    const int vB = 0;
    const float ih = 1. / (vT + 1 - vB);
    const float ic = iw * inst_cost(w, vB, vT);
    const float nic = iw * (float)(dsum(ox, vB, vT) + dsum(oy, vB, vT));
    // first-segment block: FFMA(offsets, weight, min(road, sidewalk)) in the reference SASS (:502-506)
    const float seg_g = ffma((float)(dsum(ox, vB, vT) + dsum(oy, vB, vT)), iw, seg_ground(vB, vT));
    const float seg_o = seg_object(vB, vT, ic, nic, nullptr);
    const float fn = clamp_neg(mean_of(w, m, vB, vT));
    int fni = (int)floorf(fn); fni = fni < 0 ? 0 : (fni >= D ? D - 1 : fni);
    const float data_g = w.ground_ps[vT + 1] - w.ground_ps[vB];
    const float data_o = w.lut[(size_t)fni * lstride + vT + 1] - w.lut[(size_t)fni * lstride + vB];
    const bool below = vT <= c.vhor;
    if (below) {
      float cg;
      if (c.pairwise) cg = ffma(seg_g, sw, ffma(data_g, dw, ffma(1.0f, kLn2, m.rows_log) * pw));
      else cg = ffma(seg_g, sw, ffma(ih, pw, data_g * dw));
      if (cg < w.cost[vT * 3 + GROUND]) { w.cost[vT * 3 + GROUND] = cg; w.index[vT * 3 + GROUND] = GROUND; }
    }
    float co;
    if (c.pairwise) co = ffma(seg_o, sw, ffma(data_o, dw, (((below ? kLn2 : 0.0f) + m.rows_log) + m.max_dis_log) * pw));
    else co = ffma(seg_o, sw, ffma(ih, pw, data_o * dw));
    if (co < w.cost[vT * 3 + OBJECT]) w.cost[vT * 3 + OBJECT] = co;
    w.index[vT * 3 + OBJECT] = OBJECT;
  }
  for (int vB = 1; vB < H; vB++) {
    const int pv = vB - 1;
    const bool below_prev = pv < c.vhor;
    const float cG = w.cost[pv * 3 + GROUND], cO = w.cost[pv * 3 + OBJECT], cS = w.cost[pv * 3 + SKY];
    float pc = 0.f, pm = 0.f;
    if (c.pairwise) {
      pc = ffma(lg2((float)(H - vB)), kLn2, -0.0f);
      pm = clamp_neg(mean_of(w, m, w.index[pv * 3 + OBJECT] / 3, pv));
    }
    for (int vT = vB; vT < H; vT++) {
      const float ih = 1. / (vT + 1 - vB);
      const float ic = iw * inst_cost(w, vB, vT);
      const float nic = iw * (float)(dsum(ox, vB, vT) + dsum(oy, vB, vT));
      const float seg_g = seg_ground(vB, vT) + nic;
      const float seg_o = seg_object(vB, vT, ic, nic, nullptr);
      const float seg_s = nic + (float)dsum(S + (size_t)10 * hs2, vB, vT);
      const float fn = clamp_neg(mean_of(w, m, vB, vT));
      int fni = (int)floorf(fn); fni = fni < 0 ? 0 : (fni >= D ? D - 1 : fni);
      const float data_o = w.lut[(size_t)fni * lstride + vT + 1] - w.lut[(size_t)fni * lstride + vB];
      if (below_prev) {  // ground :687-728
        const float data_g = w.ground_ps[vT + 1] - w.ground_ps[vB];
        float p1 = cG, p2 = cO, cg;
        if (c.pairwise) {
          const float prev = ffma(-kLn2, kLg2_03, pc);
          p1 = ffma(prev, pw, p1); p2 = ffma(prev, pw, p2);
          cg = ffma(seg_g, sw, ffma(data_g, dw, fminf(p1, p2) * pw));
        } else {
          cg = ffma(seg_g, sw, ffma(ih, pw, data_g * dw));
        }
        if (cg < w.cost[vT * 3 + GROUND]) {
          w.cost[vT * 3 + GROUND] = cg;
          w.index[vT * 3 + GROUND] = vB * 3 + (p1 < p2 ? GROUND : OBJECT);
        }
      } else {  // sky :729-775
        const float data_s = w.sky_ps[vT + 1] - w.sky_ps[vB];
        float p1 = cG, p2 = cO, cs;
        if (c.pairwise) {
          p1 = ffma(c.gf[pv] < 1.0f ? pc : kInf, pw, p1);
          p2 = ffma(pm < eps ? kInf : ffma(kLn2, 1.0f, pc), pw, p2);
          cs = ffma(seg_s, sw, ffma(data_s, dw, fminf(p1, p2) * pw));
        } else {
          cs = ffma(seg_s, sw, ffma(ih, pw, data_s * dw));
        }
        if (cs < w.cost[vT * 3 + SKY]) {
          w.cost[vT * 3 + SKY] = cs;
          w.index[vT * 3 + SKY] = vB * 3 + (p1 < p2 ? GROUND : OBJECT);
        }
      }
      // object :777-837
      float p1 = cG, p2 = cO, p3 = cS, co;
      if (c.pairwise) {
        {  // from ground :120-144
          const float fnp = clamp_neg(c.gf[pv]);
          float nld;
          if (fn > fnp + eps) nld = neg_log_div(m.cfg.pgrav, (-fnp + m.max_disf) - eps);
          else if (fn < fnp - eps) nld = neg_log_div(m.cfg.pblg, fnp - eps);
          else nld = neg_log_div((-m.cfg.pgrav + 1.0f) - m.cfg.pblg, eps + eps);
          p1 = ffma(nld + (pc + kNegLog07), pw, p1);
        }
        {  // from object :146-171
          int ipm = (int)pm; ipm = ipm < 0 ? 0 : (ipm >= D ? D - 1 : ipm);
          const float dd = clamp_neg(m.odr[ipm]);
          const float base = pc + (below_prev ? kNegLog07 : kLn2);
          float tr = kInf;
          if (fn > pm + dd) tr = base + neg_log_div(m.cfg.pord, -dd + (-pm + m.max_disf));
          else if (fn < pm - dd) tr = base + neg_log_div(-m.cfg.pord + 1.0f, pm - dd);
          p2 = ffma(tr, pw, p2);
        }
        {  // from sky :173-183
          float tr = kInf;
          if (fn > eps) tr = pc + ffma(lg2(m.max_disf - eps), kLn2, -0.0f);
          p3 = ffma(tr, pw, p3);
        }
        co = ffma(seg_o, sw, ffma(data_o, dw, fminf(p3, fminf(p1, p2)) * pw));
      } else {
        co = ffma(seg_o, sw, ffma(ih, pw, data_o * dw));
      }
      if (co < w.cost[vT * 3 + OBJECT]) {
        w.cost[vT * 3 + OBJECT] = co;
        int mp = p1 < p2 ? GROUND : OBJECT;
        if (p3 < fminf(p1, p2)) mp = SKY;
        w.index[vT * 3 + OBJECT] = vB * 3 + mp;
      }
    }
  }
  // ---- backtracking :843-955 ----
  int vT = H - 1;
  int type = OBJECT;
  {
    const float lg = w.cost[vT * 3 + GROUND], lo = w.cost[vT * 3 + OBJECT], ls = w.cost[vT * 3 + SKY];
    if (lg < lo) type = GROUND;
    if (ls < fminf(lg, lo)) type = SKY;
  }
  int i = 0, prev_vT;
  do {
    const int idx = vT * 3 + type;
    prev_vT = w.index[idx] / 3 - 1;
    isx_section s;
    s.vT = vT; s.type = type; s.vB = prev_vT + 1;
    s.disparity = mean_of(w, m, s.vB, s.vT);
    s.cost = fminf(w.cost[idx], 1e4);
    const float rn = rcp((float)(s.vT + 1 - s.vB));
    s.instance_meanx = (float)(w.mx[s.vT + 1] - w.mx[s.vB]) * rn;
    s.instance_meany = (float)(w.my[s.vT + 1] - w.my[s.vB]) * rn;
    if (type == GROUND) {
      s.semantic_class = ((float)dsum(S, s.vB, s.vT) < (float)dsum(S + hs2, s.vB, s.vT)) ? 0 : 1;
    } else if (type == SKY || s.disparity < 1.0f) {
      s.type = SKY; s.semantic_class = 10;
    } else {
      const float ic = iw * inst_cost(w, s.vB, s.vT);
      const float nic = iw * (float)(dsum(ox, s.vB, s.vT) + dsum(oy, s.vB, s.vT));
      seg_object(s.vB, s.vT, ic, nic, &s.semantic_class);
    }
    out[i++] = s;
    type = w.index[idx] % 3;
    vT = prev_vT;
  } while (prev_vT != -1 && i < kMaxSections - 1);
  isx_section term; std::memset(&term, 0, sizeof term); term.type = -1;
  out[i] = term;
}

}  // namespace

extern "C" {

// One frame through the restated path. sections: [realcols][200]. If cost_table /
// index_table / joined / object_lut are non-null they receive the intermediates
// ([C][H][3], [C][H][3], [C][H], [C][D][H+1]).  Returns 0, or the number of
// instances written through *n_inst.
int orc_compute_frame(const isx_config *cfg, int pairwise, const float *disparity, const int32_t *segmentation,
                      const isx_road *road, int nthreads, isx_section *sections, isx_instance *instances,
                      int inst_capacity, int *n_inst, float *cost_table, int32_t *index_table, float *joined,
                      float *object_lut) {
  Model m(*cfg);
  std::vector<float> gf, ng, ig;
  const int vhor = m.ground(*road, gf, ng, ig);
  Ctx c{m, gf, ng, ig, vhor, pairwise != 0};
  const int C = m.realcols, H = m.rows, D = m.max_dis;
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
  {
    _mm_setcsr(_mm_getcsr() | 0x8040);  // FTZ + DAZ: the reference's fp32 ops are all .ftz
    ColumnWork w;
#pragma omp for schedule(dynamic, 1)
    for (int col = 0; col < C; col++) {
      process_column(c, col, disparity, segmentation, w, sections + (size_t)col * kMaxSections);
      if (cost_table) std::memcpy(cost_table + (size_t)col * H * 3, w.cost.data(), sizeof(float) * H * 3);
      if (index_table) std::memcpy(index_table + (size_t)col * H * 3, w.index.data(), sizeof(int) * H * 3);
      if (joined) std::memcpy(joined + (size_t)col * H, w.d.data(), sizeof(float) * H);
      if (object_lut)
        for (int fn = 0; fn < D; fn++)
          std::memcpy(object_lut + ((size_t)col * D + fn) * (H + 1), &w.lut[(size_t)fn * (m.rows_p2 + 1)],
                      sizeof(float) * (H + 1));
    }
  }
  // ---- instance grouping: candidates in (class, column, index) order ----
  int total = 0;
  for (int k = 0; k < 8; k++) {
    std::vector<float> xy; std::vector<uint8_t> cand; std::vector<int> cols, idxs;
    for (int col = 0; col < C; col++)
      for (int j = 0; j < kMaxSections; j++) {
        const isx_section &s = sections[(size_t)col * kMaxSections + j];
        if (s.type == -1) break;
        if (s.type == OBJECT && s.semantic_class == 11 + k) {
          xy.push_back(s.instance_meanx); xy.push_back(s.instance_meany);
          cand.push_back((s.vT + 1 - s.vB) >= cfg->size_filter);
          cols.push_back(col); idxs.push_back(j);
        }
      }
    const int n = (int)cols.size();
    if (n == 0) continue;
    std::vector<int> labels(n);
    isx_oracle::dbscan_sizefilter(xy.data(), n, cfg->eps, cfg->min_pts, cand.data(), labels.data());
    for (int i = 0; i < n; i++, total++)
      if (instances && total < inst_capacity) instances[total] = isx_instance{cols[i], idxs[i], labels[i], 11 + k};
  }
  if (n_inst) *n_inst = total;
  return 0;
}

int orc_max_threads(void) { return omp_get_max_threads(); }

// dbscan_def.h on its own (unit tests vs sklearn).
void orc_dbscan(const float *xy, int n, float eps, int min_pts, const uint8_t *core_candidate, int *labels) {
  isx_oracle::dbscan_sizefilter(xy, n, eps, min_pts, core_candidate, labels);
}

}  // extern "C"
