/*
 * oracle/dvr_ref.c -- CPU restatement of ViDAR's voxel ray-casters.  TEST INFRASTRUCTURE:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this.  The product path (vidar_b200/) never does.
 *
 * It restates, array-for-array, what the reference CUDA kernels compute per ray:
 *   dvr.init / dvxlr.init      third_lib/dvr/dvr.cu:14-63
 *   dvr.render_forward         third_lib/dvr/dvr.cu:65-317
 *   dvr.render                 third_lib/dvr/dvr.cu:385-627
 *   dvxlr.render               third_lib/dvxlr/dvxlr.cu:160-457
 *   dvxlr.get_grad_sigma       third_lib/dvxlr/dvxlr.cu:63-112
 *   dvxlr_v2.render_v2         third_lib/dvxlr/dvxlr_v2.cu:119-427 (indicator/ray_pred :408-423)
 *   dvxlr_v2.get_grad_sigma_v2 third_lib/dvxlr/dvxlr_v2.cu:12-67
 * i.e. it keeps the per-ray path / csd / p / d / dt arrays and the backward D_i recursion of
 * the reference (the CUDA product uses a register-only closed form instead, so the two are
 * independent derivations of the same numbers).  All traversal math is double like the
 * reference.  Parity status: the reference ships no golden vectors for this path
 * (SURVEY.md section 8c); this file is pinned against the reference CUDA build itself
 * (oracle/build_ref.sh -> oracle/_ref, run on the GPU box) through tests/golden/dvr_*.npz.
 *
 * Differences from the reference that are deliberate:
 *   - gradients accumulate in double and are rounded to float once (the reference's
 *     dvr.render `+=` is a data race, dvr.cu:621-622; dvxlr uses fp32 atomics);
 *   - rays with tindex >= T (T != 1) or >= To are skipped instead of asserting / reading
 *     out of bounds; a NaN direction that never enters the grid stops after a step bound
 *     instead of looping forever.
 * Rays are independent -> OpenMP over rays.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { VAR_FORWARD = 0, VAR_RENDER = 1, VAR_DVXLR = 2 };

typedef struct {
  int N, M, T, To, Z, Y, X;
} dims_t;

typedef struct {
  int cap;
  int *px, *py, *pz;
  double *csd, *p, *d, *dt, *dd;
} scratch_t;

static void scratch_init(scratch_t* s, int cap) {
  s->cap = cap;
  s->px = (int*)malloc(sizeof(int) * cap);
  s->py = (int*)malloc(sizeof(int) * cap);
  s->pz = (int*)malloc(sizeof(int) * cap);
  s->csd = (double*)malloc(sizeof(double) * cap);
  s->p = (double*)malloc(sizeof(double) * cap);
  s->d = (double*)malloc(sizeof(double) * cap);
  s->dt = (double*)malloc(sizeof(double) * cap);
  s->dd = (double*)malloc(sizeof(double) * cap);
}
static void scratch_free(scratch_t* s) {
  free(s->px); free(s->py); free(s->pz);
  free(s->csd); free(s->p); free(s->d); free(s->dt); free(s->dd);
}

static inline int clampi(int v, int hi) { v = v < hi ? v : hi - 1; return v >= 0 ? v : 0; }

typedef struct {
  int count;      /* voxels recorded inside the grid */
  double gt_raw;  /* |end - origin| */
  double exp_d, p_out, max_d;
  int ts;
} trace_t;

/*
 * March one ray and fill the scratch arrays.  Returns 0 if the ray is skipped (padding)
 * -- in that case outputs keep their initial values, like the reference's early return.
 */
static int trace_ray(const dims_t* D, int variant, const float* sigma, const float* origin,
                     const float* points, const float* tindex, int n, int c, scratch_t* S,
                     trace_t* out) {
  const float tf = tindex[(size_t)n * D->M + c];
  if (tf < 0) return 0;
  const int t = (int)tf;
  if (t >= D->To) return 0;
  if (D->T != 1 && t >= D->T) return 0;
  const int ts = (D->T == 1) ? 0 : t;
  out->ts = ts;
  const float* o = origin + ((size_t)n * D->To + t) * 3;
  const float* e = points + ((size_t)n * D->M + c) * 3;
  const double xo = o[0], yo = o[1], zo = o[2];
  const double xe = e[0], ye = e[1], ze = e[2];
  int vx = (int)xo, vy = (int)yo, vz = (int)zo;
  double fx = (double)vx, fy = (double)vy, fz = (double)vz; /* running "path" position */
  const double rx = xe - xo, ry = ye - yo, rz = ze - zo;
  /* nvcc (default -fmad=true) contracts the reference's `rx*rx + ry*ry + rz*rz` into two fused
   * multiply-adds (order read off the SASS of the same expression); the last bit of the norm -- hence of the direction -- decides round() ties on
   * lattice-aligned rays (half-integer origins: every crossing).  Pinned by tests/golden/dvr_ties.npz
   * (the reference's own CUDA binary). */
  const double gt_d = sqrt(fma(rz, rz, fma(rx, rx, ry * ry)));     /* SASS: DMUL ry,ry ; DFMA rx,rx ; DFMA rz,rz */
  out->gt_raw = gt_d;
  const double dx = rx / gt_d, dy = ry / gt_d, dz = rz / gt_d;
  const int sx = (dx >= 0) ? 1 : -1, sy = (dy >= 0) ? 1 : -1, sz = (dz >= 0) ? 1 : -1;
  /* first boundary: render uses the standard 0:+1 rule, the two others -1:+1 */
  const int neg = (variant == VAR_RENDER) ? 0 : -1;
  const double bx = vx + (sx < 0 ? neg : 1), by = vy + (sy < 0 ? neg : 1), bz = vz + (sz < 0 ? neg : 1);
  double tmx = (dx != 0) ? (bx - xo) / dx : DBL_MAX;
  double tmy = (dy != 0) ? (by - yo) / dy : DBL_MAX;
  double tmz = (dz != 0) ? (bz - zo) / dz : DBL_MAX;
  const double tdx = (dx != 0) ? sx / dx : DBL_MAX;
  const double tdy = (dy != 0) ? sy / dy : DBL_MAX;
  const double tdz = (dz != 0) ? sz / dz : DBL_MAX;
  const int rounded = (variant != VAR_RENDER);
  const int merge = (variant == VAR_DVXLR);
  const float* sg = sigma + ((size_t)n * D->T + ts) * D->Z * (size_t)D->Y * D->X;

  int count = 0, was_inside = 0;
  double last_d = 0.0;
  long long guard = 4LL * ((long long)D->X + D->Y + D->Z) + 64 +
                    (long long)(fabs(xo) + fabs(yo) + fabs(zo)) * 2;
  if (!(guard < (1LL << 24))) guard = 1LL << 24;
  while (guard-- > 0) {
    const int inside = (0 <= vx && vx < D->X) && (0 <= vy && vy < D->Y) && (0 <= vz && vz < D->Z);
    if (inside) {
      was_inside = 1;
      if (count >= S->cap) break; /* cannot happen for cap >= X+Y+Z+3 */
      if (rounded) {
        S->px[count] = clampi((int)round(fx), D->X);
        S->py[count] = clampi((int)round(fy), D->Y);
        S->pz[count] = clampi((int)round(fz), D->Z);
      } else {
        S->px[count] = vx; S->py[count] = vy; S->pz[count] = vz;
      }
    } else if (was_inside) {
      break;
    } else if (last_d > gt_d) {
      break;
    }
    double dcur;
    if (tmx < tmy) {
      if (tmx < tmz) { dcur = tmx; vx += sx; tmx += tdx; }
      else           { dcur = tmz; vz += sz; tmz += tdz; }
    } else {
      if (tmy < tmz) { dcur = tmy; vy += sy; tmy += tdy; }
      else           { dcur = tmz; vz += sz; tmz += tdz; }
    }
    if (rounded) {
      /* nvcc's default -fmad=true contracts the reference's `path_v += max(0,_d-last_d)*d`
       * (dvr.cu:253-255) into one DFMA.  The single rounding matters: an origin with a .5
       * fractional part puts every crossing of that axis exactly on a round() tie, so the
       * recorded voxel depends on the last ulp.  fma() reproduces the reference binary. */
      const double adv = fmax(0.0, dcur - last_d);
      fx = fma(adv, dx, fx); fy = fma(adv, dy, fy); fz = fma(adv, dz, fz);
    }
    if (inside) {
      const double s = sg[((size_t)S->pz[count] * D->Y + S->py[count]) * D->X + S->px[count]];
      if (merge && count >= 1 && S->px[count - 1] == S->px[count] &&
          S->py[count - 1] == S->py[count] && S->pz[count - 1] == S->pz[count]) {
        count--;
        last_d -= S->dt[count];
      }
      const double delta = fmax(0.0, dcur - last_d);
      const double sd = s * delta;
      if (count == 0) {
        S->csd[0] = sd;
        S->p[0] = 1 - exp(-sd);
      } else {
        S->csd[count] = S->csd[count - 1] + sd;
        S->p[count] = exp(-S->csd[count - 1]) - exp(-S->csd[count]);
      }
      S->d[count] = dcur;
      S->dt[count] = delta;
      count++;
    }
    last_d = dcur;
  }
  out->count = count;
  if (count > 0) {
    double acc = 0.0;
    for (int i = 0; i < count; i++) acc += S->p[i] * S->d[i];
    out->p_out = exp(-S->csd[count - 1]);
    out->max_d = S->d[count - 1];
    out->exp_d = acc + out->p_out * out->max_d;
  }
  return 1;
}

/* d exp_d / d sigma_i for every recorded voxel (the reference's three backward loops) */
static void backward_lists(const scratch_t* S, const trace_t* tr) {
  const int count = tr->count;
  for (int i = count - 1; i >= 0; i--) {
    if (i == count - 1) S->dd[i] = tr->p_out * tr->max_d;
    else S->dd[i] = S->dd[i + 1] - exp(-S->csd[i]) * (S->d[i + 1] - S->d[i]);
  }
  for (int i = count - 1; i >= 0; i--) S->dd[i] *= S->dt[i];
  for (int i = count - 1; i >= 0; i--) S->dd[i] -= S->dt[i] * tr->p_out * tr->max_d;
}

static int scratch_cap(const dims_t* D) { return D->X + D->Y + D->Z + 8; }

void oracle_dvr_init(const float* points, const float* tindex, float* occupancy, int N, int M,
                     int T, int Z, int Y, int X) {
  for (int n = 0; n < N; n++)
    for (int c = 0; c < M; c++) {
      const float tf = tindex[(size_t)n * M + c];
      if (tf < 0) continue;
      const int t = (int)tf;
      if (T != 1 && t >= T) continue;
      const int ts = (T == 1) ? 0 : t;
      const float* e = points + ((size_t)n * M + c) * 3;
      const int vx = (int)e[0], vy = (int)e[1], vz = (int)e[2];
      if (0 <= vx && vx < X && 0 <= vy && vy < Y && 0 <= vz && vz < Z)
        occupancy[((((size_t)n * T + ts) * Z + vz) * Y + vy) * X + vx] = 1.f;
    }
}

/* pred_dist / gt_dist must be pre-filled with -1 by the caller. */
void oracle_dvr_render_forward(const float* sigma, const float* origin, const float* points,
                               const float* tindex, float* pred_dist, float* gt_dist, int N,
                               int M, int T, int To, int Z, int Y, int X, int train_phase) {
  const dims_t D = {N, M, T, To, Z, Y, X};
#pragma omp parallel
  {
    scratch_t S;
    scratch_init(&S, scratch_cap(&D));
#pragma omp for schedule(dynamic, 64)
    for (long long r = 0; r < (long long)N * M; r++) {
      const int n = (int)(r / M), c = (int)(r % M);
      trace_t tr;
      if (!trace_ray(&D, VAR_FORWARD, sigma, origin, points, tindex, n, c, &S, &tr)) continue;
      if (tr.count == 0) continue;
      double gt = tr.gt_raw;
      if (train_phase == 1) gt = fmin(gt, tr.max_d);
      pred_dist[r] = (float)tr.exp_d;
      gt_dist[r] = (float)gt;
    }
    scratch_free(&S);
  }
}

/* forward of dvxlr.render without the lists (used as the fast oracle for the fused op) */
void oracle_dvxlr_forward(const float* sigma, const float* origin, const float* points,
                          const float* tindex, float* pred_dist, float* gt_dist, int N, int M,
                          int T, int To, int Z, int Y, int X) {
  const dims_t D = {N, M, T, To, Z, Y, X};
#pragma omp parallel
  {
    scratch_t S;
    scratch_init(&S, scratch_cap(&D));
#pragma omp for schedule(dynamic, 64)
    for (long long r = 0; r < (long long)N * M; r++) {
      const int n = (int)(r / M), c = (int)(r % M);
      trace_t tr;
      if (!trace_ray(&D, VAR_DVXLR, sigma, origin, points, tindex, n, c, &S, &tr)) continue;
      if (tr.count == 0) continue;
      pred_dist[r] = (float)tr.exp_d;
      gt_dist[r] = (float)fmin(tr.gt_raw, tr.max_d);
    }
    scratch_free(&S);
  }
}

/* loss_type: 0 l1 (and "bce"), 1 l2, 2 absrel.  grad_sigma_out fully written. */
void oracle_dvr_render(const float* sigma, const float* origin, const float* points,
                       const float* tindex, float* pred_dist, float* gt_dist,
                       float* grad_sigma_out, int N, int M, int T, int To, int Z, int Y, int X,
                       int loss_type) {
  const dims_t D = {N, M, T, To, Z, Y, X};
  const size_t vol = (size_t)Z * Y * X, total = (size_t)N * T * vol;
  double* acc = (double*)calloc(total, sizeof(double));
#pragma omp parallel
  {
    scratch_t S;
    scratch_init(&S, scratch_cap(&D));
#pragma omp for schedule(dynamic, 64)
    for (long long r = 0; r < (long long)N * M; r++) {
      const int n = (int)(r / M), c = (int)(r % M);
      trace_t tr;
      if (!trace_ray(&D, VAR_RENDER, sigma, origin, points, tindex, n, c, &S, &tr)) continue;
      if (tr.count == 0) continue;
      const double gt = fmin(tr.gt_raw, tr.max_d);
      pred_dist[r] = (float)tr.exp_d;
      gt_dist[r] = (float)gt;
      backward_lists(&S, &tr);
      double dl = 1.0;
      if (loss_type == 0) dl = (tr.exp_d >= gt) ? 1 : -1;
      else if (loss_type == 1) dl = tr.exp_d - gt;
      else if (loss_type == 2) dl = (tr.exp_d >= gt) ? (1.0 / gt) : -(1.0 / gt);
      double* g = acc + ((size_t)n * T + tr.ts) * vol;
      for (int i = 0; i < tr.count; i++) {
        const size_t vo = ((size_t)S.pz[i] * Y + S.py[i]) * X + S.px[i];
#pragma omp atomic
        g[vo] += dl * S.dd[i];
      }
    }
    scratch_free(&S);
  }
  for (size_t i = 0; i < total; i++) grad_sigma_out[i] = (float)acc[i];
  free(acc);
}

/*
 * dvxlr.render (sigma_regul/ray_pred/indicator NULL) and dvxlr_v2.render_v2.
 * Caller pre-fills: pred/gt = -1, dd_dsigma/indices/ray_pred = 0, indicator = -1.
 */
void oracle_dvxlr_render(const float* sigma, const float* origin, const float* points,
                         const float* tindex, const float* sigma_regul, float* pred_dist,
                         float* gt_dist, float* dd_dsigma, float* indices, float* ray_pred,
                         float* indicator, int N, int M, int T, int To, int Z, int Y, int X,
                         int max_d) {
  const dims_t D = {N, M, T, To, Z, Y, X};
  const size_t vol = (size_t)Z * Y * X;
#pragma omp parallel
  {
    scratch_t S;
    scratch_init(&S, scratch_cap(&D));
#pragma omp for schedule(dynamic, 64)
    for (long long r = 0; r < (long long)N * M; r++) {
      const int n = (int)(r / M), c = (int)(r % M);
      trace_t tr;
      if (!trace_ray(&D, VAR_DVXLR, sigma, origin, points, tindex, n, c, &S, &tr)) continue;
      if (tr.count == 0) continue;
      pred_dist[r] = (float)tr.exp_d;
      gt_dist[r] = (float)fmin(tr.gt_raw, tr.max_d);
      backward_lists(&S, &tr);
      int reached = 0;
      for (int i = 0; i < tr.count && i < max_d; i++) {
        dd_dsigma[(size_t)r * max_d + i] = (float)S.dd[i];
        float* id = indices + ((size_t)r * max_d + i) * 3;
        id[0] = (float)S.pz[i]; id[1] = (float)S.py[i]; id[2] = (float)S.px[i];
        if (sigma_regul) {
          float ind = 0.f;
          if (!reached && S.d[i] >= tr.gt_raw) { ind = 1.f; reached = 1; }
          indicator[(size_t)r * max_d + i] = ind;
          ray_pred[(size_t)r * max_d + i] =
              sigma_regul[((size_t)n * T + tr.ts) * vol + ((size_t)S.pz[i] * Y + S.py[i]) * X + S.px[i]];
        }
      }
    }
    scratch_free(&S);
  }
}

/*
 * dvxlr.get_grad_sigma / get_grad_sigma_v2: scatter every list slot (all max_d of them,
 * padding included, like the reference) into the volume.  Double accumulation.
 */
void oracle_dvxlr_get_grad_sigma(const float* elementwise_mult, const float* indices,
                                 const float* tindex, const float* indicator,
                                 const float* grad_ray_pred, float* grad_sigma,
                                 float* grad_sigma_regul, int N, int M, int T, int Z, int Y,
                                 int X, int max_d) {
  const size_t vol = (size_t)Z * Y * X, total = (size_t)N * T * vol;
  double* a = (double*)calloc(total, sizeof(double));
  double* b = indicator ? (double*)calloc(total, sizeof(double)) : NULL;
  for (int n = 0; n < N; n++)
    for (int c = 0; c < M; c++) {
      const size_t r = (size_t)n * M + c;
      const float tf = tindex[r];
      if (tf < 0) continue;
      const int t = (int)tf;
      if (T != 1 && t >= T) continue;
      const int ts = (T == 1) ? 0 : t;
      for (int i = 0; i < max_d; i++) {
        const float* id = indices + (r * max_d + i) * 3;
        const int z = (int)id[0], y = (int)id[1], x = (int)id[2];
        const size_t vo = ((size_t)n * T + ts) * vol + ((size_t)z * Y + y) * X + x;
        a[vo] += elementwise_mult[r * max_d + i];
        if (indicator && indicator[r * max_d + i] >= 0) b[vo] += grad_ray_pred[r * max_d + i];
      }
    }
  for (size_t i = 0; i < total; i++) grad_sigma[i] = (float)a[i];
  if (b) for (size_t i = 0; i < total; i++) grad_sigma_regul[i] = (float)b[i];
  free(a);
  free(b);
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
