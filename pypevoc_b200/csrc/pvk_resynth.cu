// pvk_resynth.cu -- phase-preserving sum-of-sinusoids resynthesis for sm_100a
// (SinSum.synth -> RegPartial.synth, PVAnalysis.py:1053-1070,684-756).
//
// One CTA owns one output block of `hop` samples.  Every hop-block of a partial depends on
// at most 4 neighbouring f, 3 mag and 2 realph values of that partial and the reference's
// cumsum never crosses a block (:703-736), so blocks are independent.
// The CTA walks the frame rows that can sound in its block (the block's own row for
// partial bodies, +-ceil(E/hop) rows for fade-in heads / fade-out tails), one thread per
// row slot turns the partial's neighbourhood into closed-form coefficients staged in shared
// memory -- the phase of a block is a piecewise quadratic of the sample index because the
// frequency is piecewise linear with one knot per block -- and then every thread renders
// its samples: fp64 phase polynomial in cycles, exact range reduction, fp32 cosine.
// Not HBM bound (4-8 bytes written per output sample, ~0.03-0.1 B per partial-sample): the
// binding resources are the FP64 pipe and MUFU; see DESIGN.md.
#include "pvk_common.cuh"

namespace pvk {

struct RParams {
  const int32_t *tid;       // [F, K]
  const int32_t *tstart, *tlen;
  const int64_t *toff;
  const double *pf, *pmag, *prealph;
  int64_t F;
  int K;
  double sr, fstep, dfr;
  int h;                    // synthesis hop
  int E;                    // int(dfr*hop*edge)  (:740, :1056)
  int dE;                   // ceil(E / h)
  int minframes;
  double *out;
  int64_t nout, block0;
};

constexpr int R_S = 8;      // samples per thread per pass

// np.interp(n, h*(a + arange(nfr)), v) for integer n >= 0 (:701-702), clamped at both ends
__device__ __forceinline__ double interp_at(const double *__restrict__ v, int nfr, double a, int h, double n) {
  const double kf = n / (double)h - a;
  if (!(kf > 0.0)) return v[0];
  if (kf >= (double)(nfr - 1)) return v[nfr - 1];
  int j = (int)floor(kf);
  if (j > nfr - 2) j = nfr - 2;
  const double xj = (double)h * (a + (double)j);
  const double slope = (v[j + 1] - v[j]) / (double)h;
  return slope * (n - xj) + v[j];
}

// value(q) on block ii, q in [0,h): q < qk ? c0 + s0*q : c1 + s1*q
struct Lin2 { double c0, s0, c1, s1; int qk; };

__device__ __forceinline__ Lin2 interp_block(const double *__restrict__ v, int nfr, int ii, double a, int h) {
  Lin2 L;
  const double A = floor(a), alpha = a - A;
  const int k0 = ii - (int)A;                    // knot k0 sits at q = alpha*h inside this block
  const double qa = alpha * (double)h;
  L.qk = (int)ceil(qa);
  // segment 0: between knots k0-1 and k0
  if (k0 <= 0) { L.c0 = v[0]; L.s0 = 0.0; }
  else if (k0 - 1 >= nfr - 1) { L.c0 = v[nfr - 1]; L.s0 = 0.0; }
  else {
    const double slope = (v[k0] - v[k0 - 1]) / (double)h;
    L.c0 = v[k0 - 1] + slope * ((double)h - qa); L.s0 = slope;
  }
  // segment 1: between knots k0 and k0+1
  if (k0 < 0) { L.c1 = v[0]; L.s1 = 0.0; }
  else if (k0 >= nfr - 1) { L.c1 = v[nfr - 1]; L.s1 = 0.0; }
  else {
    const double slope = (v[k0 + 1] - v[k0]) / (double)h;
    L.c1 = v[k0] - slope * qa; L.s1 = slope;
  }
  return L;
}

// sum_{u<q} F(u)/sr of the piecewise linear F (cycles)
__device__ __forceinline__ double cum_cycles(const Lin2 &L, double sr, int q) {
  const double qd = (double)q, qk = (double)L.qk;
  if (q <= L.qk) return (L.c0 * qd + L.s0 * (qd * (qd - 1.0)) * 0.5) / sr;
  const double s0 = (L.c0 * qk + L.s0 * (qk * (qk - 1.0)) * 0.5) / sr;
  return s0 + (L.c1 * (qd - qk) + L.s1 * ((qd * (qd - 1.0)) - (qk * (qk - 1.0))) * 0.5) / sr;
}

// one sinusoid segment that sounds in the block: phase polynomial (cycles) and linear
// amplitude, each with one knot; 96 bytes, read back with six 128-bit shared loads
struct __align__(16) Item {
  double A0, B0, C0, A1, B1, C1;     // theta(q) = A + B q + C q^2, set 1 for q >= qk
  float m0c, m0s, m1c, m1s;          // amp(q)   = c + s q,         set 1 for q >= qkm
  int qk, qkm, qa, qb;               // knots and valid sample range [qa, qb)
  int eoff; float einv; int type; int pad;   // fade envelope: cos(pi*(q+eoff)*einv)
};

constexpr int IT_NONE = 0, IT_BODY = 1, IT_HEAD = 2, IT_TAIL = 3;
constexpr double INV_2PI = 0.15915494309189535;
constexpr double TWO_PI = 6.283185307179586;
constexpr double PI_D = 3.141592653589793;

// turn slot c of frame row r into the segment that sounds in block b (type IT_NONE if none)
__device__ __forceinline__ void make_item(const RParams &p, int64_t b, int64_t r, int c, Item &it) {
  const int h = p.h, E = p.E;
  const double sr = p.sr;
  it.type = IT_NONE;
  const int v = p.tid[r * p.K + c];
  if (v < 0) return;
  const int nfr = p.tlen[v];
  if (nfr < p.minframes) return;                                   // :1061
  const int s = p.tstart[v];
  const int ii = (int)(r - s);
  int type = IT_NONE;
  if (r == b) type = IT_BODY;
  else if (r > b && ii == 0) type = IT_HEAD;
  else if (r < b && ii == nfr - 1) type = IT_TAIL;
  if (type == IT_NONE) return;
  const int64_t o = p.toff[v];
  const double *tf = p.pf + o, *tm = p.pmag + o, *tr = p.prealph + o;
  const double af = p.dfr + 0.5, am = p.dfr;                        // knot offsets of :701 / :702
  if (type == IT_BODY) {
    const Lin2 Lf = interp_block(tf, nfr, ii, af, h);
    const double fb0 = interp_at(tf, nfr, af, h, (double)h * ii);
    const double fb1 = interp_at(tf, nfr, af, h, (double)h * (ii + 1));
    const double phcor = PI_D * (fb1 - fb0) / p.fstep / 2.0;        // :715
    const double th0 = (tr[ii] + phcor) * INV_2PI;                  // :721
    double dphc = 0.0;
    if (ii < nfr - 1) {
      const double fb2 = interp_at(tf, nfr, af, h, (double)h * (ii + 2));
      const double phcornext = PI_D * (fb2 - fb1) / p.fstep / 2.0;  // :717-718
      const double phend = TWO_PI * cum_cycles(Lf, sr, h - 1) + (tr[ii] + phcor) + TWO_PI * fb1 / sr;   // :726
      double mm = fmod(tr[ii + 1] + phcornext - phend + PI_D, TWO_PI);   // np.mod :727
      if (mm < 0.0) mm += TWO_PI;
      dphc = (mm - PI_D) * INV_2PI / (double)h;                     // :728-729, per sample, cycles
    }
    const double qk = (double)Lf.qk;
    it.A0 = th0;
    it.B0 = (Lf.c0 - 0.5 * Lf.s0) / sr + dphc;
    it.C0 = 0.5 * Lf.s0 / sr;
    it.A1 = th0 + (Lf.c0 * qk + Lf.s0 * (qk * (qk - 1.0)) * 0.5) / sr
                - (Lf.c1 * qk + Lf.s1 * (qk * (qk - 1.0)) * 0.5) / sr;
    it.B1 = (Lf.c1 - 0.5 * Lf.s1) / sr + dphc;
    it.C1 = 0.5 * Lf.s1 / sr;
    it.qk = Lf.qk;
    const Lin2 Lm = interp_block(tm, nfr, ii, am, h);
    it.m0c = (float)Lm.c0; it.m0s = (float)Lm.s0;
    it.m1c = (float)Lm.c1; it.m1s = (float)Lm.s1;
    it.qkm = Lm.qk;
    it.qa = 0; it.qb = h;
    it.eoff = 0; it.einv = 0.f;
  } else if (type == IT_HEAD) {
    // head sample qh = q + off, off = (b - s)*h + E; valid 0 <= qh < E   (:740-745)
    const int64_t off = (b - (int64_t)s) * h + E;
    const double fc = tf[0] / sr;
    const int qa = (int)(off < 0 ? -off : 0);
    int64_t qb = (int64_t)E - off;
    if (qb > h) qb = h;
    if (qb <= qa) return;
    it.A0 = tr[0] * INV_2PI - fc * (double)((int64_t)E - off);
    it.B0 = fc; it.C0 = 0.0;
    it.A1 = it.A0; it.B1 = fc; it.C1 = 0.0;
    it.qk = h;
    const float m0 = (float)interp_at(tm, nfr, am, h, 0.0);         // msig[0] :742
    it.m0c = m0; it.m0s = 0.f; it.m1c = m0; it.m1s = 0.f; it.qkm = h;
    it.qa = qa; it.qb = (int)qb;
    it.eoff = (int)off; it.einv = (float)(1.0 / (double)E);
  } else {
    // tail sample qt = q + off, off = (b - s - nfr)*h; valid 0 <= qt < E   (:748-751)
    const int64_t off = (b - (int64_t)s - nfr) * h;
    const int il = nfr - 1;
    int64_t qb = (int64_t)E - off;
    if (qb > h) qb = h;
    if (qb <= 0) return;
    const Lin2 Lf = interp_block(tf, nfr, il, af, h);
    const double fb0 = interp_at(tf, nfr, af, h, (double)h * il);
    const double fb1 = interp_at(tf, nfr, af, h, (double)h * (il + 1));
    const double phcor = PI_D * (fb1 - fb0) / p.fstep / 2.0;
    const double thl = (tr[il] + phcor) * INV_2PI + cum_cycles(Lf, sr, h - 1);   // ph[-1] :750
    const double fc = tf[il] / sr;
    it.A0 = thl + fc * (double)(off + 1);
    it.B0 = fc; it.C0 = 0.0;
    it.A1 = it.A0; it.B1 = fc; it.C1 = 0.0;
    it.qk = h;
    const float m0 = (float)interp_at(tm, nfr, am, h, (double)h * nfr);   // msig[hop*nfr] :748
    it.m0c = m0; it.m0s = 0.f; it.m1c = m0; it.m1s = 0.f; it.qkm = h;
    it.qa = 0; it.qb = (int)qb;
    it.eoff = (int)off; it.einv = (float)(1.0 / (double)E);
  }
  it.type = type;
  it.pad = 0;
}

// Thread t renders the R_S consecutive samples q = qs .. qs+R_S-1 of the block.  Fast path (a
// partial body whose phase / amplitude knots do not fall inside the thread's samples, i.e.
// practically always): the phase polynomial is evaluated once in fp64 at qs, reduced to
// [-0.5, 0.5] cycles, and advanced over the R_S samples as an fp32 quadratic in radians
// (|increment| < 4 cycles, so fp32 keeps ~3e-7 cycles); one MUFU cosine per partial-sample.
// Slow path (fade-in / fade-out segments, knots inside the chunk): fp64 per sample.
__global__ void __launch_bounds__(256) resynth_kernel(RParams p) {
  PVK_SMEM(smem);
  Item *items = reinterpret_cast<Item *>(smem);
  int *wsum = reinterpret_cast<int *>(smem + (size_t)p.K * sizeof(Item));   // 2 x 8 warp sums
  const int tid = threadIdx.x, BD = blockDim.x, NWARP = BD >> 5;
  const int lane = tid & 31, warp = tid >> 5;
  const int64_t b = p.block0 + blockIdx.x;
  const int h = p.h, K = p.K;
  const int64_t nbase = b * (int64_t)h;
  const float TWO_PI_F = 6.283185307179586f;
  int round = 0;

  for (int q0 = 0; q0 < h; q0 += R_S * BD) {
    const int qs = q0 + tid * R_S;
    const double qsd = (double)qs;
    double tot[R_S];
#pragma unroll
    for (int m = 0; m < R_S; ++m) tot[m] = 0.0;

    for (int64_t r = b - p.dE; r <= b + p.dE; ++r) {
      if (r < 0 || r >= p.F) continue;                       // uniform
      // ---- stage the segments of row r, compacted in slot order (deterministic sum order)
      int nit = 0;
      for (int c0 = 0; c0 < K; c0 += BD, ++round) {
        const int c = c0 + tid;
        Item it;
        it.type = IT_NONE;
        if (c < K) make_item(p, b, r, c, it);
        const bool on = it.type != IT_NONE;
        const unsigned mk = __ballot_sync(FULL, on);
        int *ws = wsum + (round & 1) * 8;
        if (lane == 0) ws[warp] = __popc(mk);
        __syncthreads();
        int wb = 0, tt = 0;
        for (int w = 0; w < NWARP; ++w) { const int x = ws[w]; wb += (w < warp) ? x : 0; tt += x; }
        if (on) items[nit + wb + __popc(mk & lanemask_lt())] = it;
        nit += tt;
      }
      if (nit == 0) continue;                                // uniform: nothing of this row sounds here
      __syncthreads();
      // ---- render: every thread adds all segments of this row to its samples
      float acc[R_S];
#pragma unroll
      for (int m = 0; m < R_S; ++m) acc[m] = 0.f;
      for (int c = 0; c < nit; ++c) {
        const double2 *pd = reinterpret_cast<const double2 *>(&items[c]);
        const double2 d0 = pd[0], d1 = pd[1], d2 = pd[2];    // A0 B0 | C0 A1 | B1 C1
        const float4 mf = *reinterpret_cast<const float4 *>(&items[c].m0c);
        const int4 qi = *reinterpret_cast<const int4 *>(&items[c].qk);      // qk qkm qa qb
        const int4 ei = *reinterpret_cast<const int4 *>(&items[c].eoff);    // eoff einv type pad
        const int type = ei.z;
        const bool fast = (type == IT_BODY) && (qs + R_S <= h) &&
                          (qs + R_S <= qi.x || qs >= qi.x) && (qs + R_S <= qi.y || qs >= qi.y);
        if (fast) {
          const bool s1 = qs >= qi.x;
          const double A = s1 ? d1.y : d0.x, B = s1 ? d2.x : d0.y, C = s1 ? d2.y : d1.x;
          double th = fma(fma(C, qsd, B), qsd, A);           // cycles at the first sample
          th -= rint(th);
          const double bq = fma(2.0 * C, qsd, B);            // d theta / dq at qs (cycles per sample)
          const float t0 = TWO_PI_F * (float)th, t1 = TWO_PI_F * (float)bq, t2 = TWO_PI_F * (float)C;
          const bool a1 = qs >= qi.y;
          const float as = a1 ? mf.w : mf.y;
          const float a0 = fmaf(as, (float)qs, a1 ? mf.z : mf.x);
#pragma unroll
          for (int m = 0; m < R_S; ++m) {
            const float fm = (float)m;
            const float cs = __cosf(fmaf(fmaf(t2, fm, t1), fm, t0));
            acc[m] = fmaf(fmaf(as, fm, a0), cs, acc[m]);
          }
        } else {
          const float einv = __int_as_float(ei.y);
#pragma unroll
          for (int m = 0; m < R_S; ++m) {
            const int q = qs + m;
            if (q >= qi.z && q < qi.w) {
              const double qd = (double)q;
              const bool s1 = q >= qi.x;
              const double th = fma(fma(s1 ? d2.y : d1.x, qd, s1 ? d2.x : d0.y), qd, s1 ? d1.y : d0.x);
              const float fr = (float)(th - rint(th));       // exact range reduction, [-0.5, 0.5]
              const float cs = __cosf(TWO_PI_F * fr);
              const float qf = (float)q;
              float am = (q >= qi.y) ? fmaf(mf.w, qf, mf.z) : fmaf(mf.y, qf, mf.x);
              if (type != IT_BODY) {
                const float ce = __cosf(3.14159265358979f * (float)(q + ei.x) * einv);
                am *= (type == IT_HEAD) ? 0.5f * (1.f - ce) : 0.5f * (1.f + ce);
              }
              acc[m] = fmaf(am, cs, acc[m]);
            }
          }
        }
      }
#pragma unroll
      for (int m = 0; m < R_S; ++m) tot[m] += (double)acc[m];
      __syncthreads();
    }
#pragma unroll
    for (int m = 0; m < R_S; ++m) {
      const int q = qs + m;
      const int64_t n = nbase + q;
      if (q < h && n < p.nout) p.out[n - p.block0 * (int64_t)h] = tot[m];
    }
  }
}

}  // namespace pvk

using namespace pvk;

extern "C" int pvk_resynth(const int32_t *tid, int64_t nframes, int npks, const int32_t *tstart,
                           const int32_t *tlen, const int64_t *toff, const double *pf, const double *pmag,
                           const double *prealph, double sr, int hop, int nfft, int hop_an, double edge,
                           int minframes, double *out, int64_t nout, int64_t block0, int64_t nblocks,
                           void *stream) {
  PVK_REQUIRE(hop >= 1 && nfft >= 1 && hop_an >= 1, "pvk_resynth: hop=%d nfft=%d hop_an=%d must be >= 1", hop, nfft, hop_an);
  PVK_REQUIRE(npks >= 1 && npks <= PVK_MAX_NPKS, "pvk_resynth: npks=%d must be in [1, %d]", npks, PVK_MAX_NPKS);
  PVK_REQUIRE(sr > 0.0 && edge >= 0.0, "pvk_resynth: sr and edge must be positive");
  PVK_REQUIRE(nout >= 0 && nframes >= 0 && block0 >= 0, "pvk_resynth: negative sizes");
  const int64_t nblk_total = (nout + hop - 1) / hop;
  if (nblocks < 0) nblocks = nblk_total - block0;
  if (nblocks <= 0 || nout == 0) return PVK_OK;
  PVK_REQUIRE(block0 + nblocks <= nblk_total, "pvk_resynth: block range [%lld, %lld) exceeds %lld blocks",
              (long long)block0, (long long)(block0 + nblocks), (long long)nblk_total);
  PVK_REQUIRE(out != nullptr, "pvk_resynth: out is NULL");
  PVK_REQUIRE(nframes == 0 || (tid && tstart && tlen && toff && pf && pmag && prealph),
              "pvk_resynth: NULL pointer argument");
  RParams p;
  p.tid = tid; p.tstart = tstart; p.tlen = tlen; p.toff = toff;
  p.pf = pf; p.pmag = pmag; p.prealph = prealph;
  p.F = nframes; p.K = npks; p.sr = sr;
  p.fstep = sr / (double)nfft;                               // :825
  const double overlap = (double)hop_an / (double)nfft;      // :824
  p.dfr = 1.0 / overlap / 2.0;                               // :687
  p.h = hop;
  p.E = (int)(p.dfr * (double)hop * edge);                   // :740
  p.dE = (p.E + hop - 1) / hop;
  p.minframes = minframes;
  p.out = out; p.nout = nout; p.block0 = block0;
  int bd = ((hop + R_S - 1) / R_S + 31) / 32 * 32;
  if (bd > 256) bd = 256;
  if (bd < 32) bd = 32;
  const int smem = npks * (int)sizeof(Item) + 64;
  if (smem > 48 * 1024) {
    if (PVK_SET_SMEM(resynth_kernel, smem) != 0) {
      set_error("pvk_resynth: cannot reserve %d bytes of shared memory", smem);
      return PVK_ERR_CUDA;
    }
  }
  PVK_REQUIRE(nblocks < (int64_t)2147483647, "pvk_resynth: too many blocks");
  PVK_LAUNCH(resynth_kernel, dim3((unsigned)nblocks), dim3(bd), smem, stream, p);
  PVK_CHECK_LAUNCH("pvk_resynth");
  return PVK_OK;
}
