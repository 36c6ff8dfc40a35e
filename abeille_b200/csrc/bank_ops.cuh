/* bank_ops.cuh -- everything that happens to a bank between two transport() calls, on the device.
 *
 *   place_sites          deterministic fission-bank order without a sort (see transport.cuh)
 *   weight stats/scale   PowerIterator::normalize_weights            src/power_iterator.cpp:538-586
 *   to_particles         bank rebuild with fresh history ids         src/power_iterator.cpp:386-404
 *   entropy binning      Entropy::add_point                          src/entropy.cpp:32-60
 *   score_source         SourceMeshTally::score_source               src/source_mesh_tally.cpp:30-78
 *   cancellation         ApproximateMeshCancelator                   src/approximate_mesh_cancelator.cpp:97-190
 *   source sampling      Simulation::sample_sources, Source::generate_particle
 *                                                                    src/simulation.cpp:55-77, src/source.cpp:44-90
 *   record_generation    MeshTally::record_generation                src/mesh_tally.cpp:121-150
 * All of these are streaming kernels bounded by HBM bandwidth; grids are sized in multiples of the SM count.
 */
#pragma once
#include "transport.cuh"

namespace abl {

// ---- exclusive scan of per-history site counts (three passes over 4096-element tiles) --------------------
#define ABL_SCAN_THREADS 1024
#define ABL_SCAN_ITEMS 4
#define ABL_SCAN_TILE (ABL_SCAN_THREADS * ABL_SCAN_ITEMS)

__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sums[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = warp_sums[lane];
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += n;
    }
    warp_sums[lane] = winc - w;  // exclusive
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  const uint32_t r = warp_sums[wid] + inc - v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(ABL_SCAN_THREADS) scan_tile_sums_kernel(const uint32_t* __restrict__ in, uint64_t n,
                                                                           uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t total;
  const uint64_t base = (uint64_t)blockIdx.x * ABL_SCAN_TILE + (uint64_t)threadIdx.x * ABL_SCAN_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < ABL_SCAN_ITEMS; k++)
    if (base + k < n) s += in[base + k];
  (void)block_exclusive_scan_1024(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(ABL_SCAN_THREADS) scan_top_kernel(uint32_t* tile_sums, uint32_t ntiles, uint32_t* grand_total) {
  __shared__ uint32_t total;
  uint32_t carry = 0;
  for (uint32_t base = 0; base < ntiles; base += ABL_SCAN_THREADS) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < ntiles ? tile_sums[i] : 0;
    const uint32_t ex = block_exclusive_scan_1024(v, &total);
    if (i < ntiles) tile_sums[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(ABL_SCAN_THREADS) scan_apply_kernel(const uint32_t* __restrict__ in, uint64_t n,
                                                                       const uint32_t* __restrict__ tile_offsets,
                                                                       uint32_t* __restrict__ out) {
  __shared__ uint32_t total;
  const uint64_t base = (uint64_t)blockIdx.x * ABL_SCAN_TILE + (uint64_t)threadIdx.x * ABL_SCAN_ITEMS;
  uint32_t v[ABL_SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < ABL_SCAN_ITEMS; k++) {
    v[k] = base + k < n ? in[base + k] : 0;
    s += v[k];
  }
  uint32_t ex = block_exclusive_scan_1024(s, &total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < ABL_SCAN_ITEMS; k++) {
    if (base + k < n) out[base + k] = ex;
    ex += v[k];
  }
}

// ---- fission bank placement ---------------------------------------------------------------------------------
// Scratch sites arrive in random parent order; their place in the bank is offsets[parent] + rank.  Writing them there directly
// costs eleven scattered 8-byte stores per site (11 read-modify-written 32 B sectors: 4.3 ms for 1e7 sites, 6 % of the HBM
// rate).  Instead: (1) one scattered 4-byte store per site builds the inverse map row -> scratch index, (2) the rows are
// written in order -- every store of a warp is one coalesced line -- and the only scattered access left is the read of the
// 80-byte record (three sectors).
// did: daughter ids of the scratch sites when they differ from the rank carried by the site (noise mode), else null
__global__ void __launch_bounds__(256) site_inverse_kernel(const Site* __restrict__ sites, uint64_t n_sites,
                                                           const uint32_t* __restrict__ offsets, uint64_t capacity,
                                                           uint32_t* __restrict__ inv) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_sites; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long tag = *reinterpret_cast<const unsigned long long*>(reinterpret_cast<const char*>(sites + i) + 72);
    const uint32_t parent = (uint32_t)(tag & 0xffffffffULL), rank = (uint32_t)(tag >> 32);
    const uint64_t pos = (uint64_t)offsets[parent] + rank;
    if (pos < capacity) inv[pos] = (uint32_t)i;  // capacity overflow is reported by the host from the total count
  }
}
__global__ void __launch_bounds__(256) place_sites_kernel(const Site* __restrict__ sites, uint64_t n_sites,
                                                          const uint32_t* __restrict__ inv, BankView in, BankView out,
                                                          const uint32_t* __restrict__ did) {
  for (uint64_t pos = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; pos < n_sites; pos += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t i = inv[pos];
    const double2* src = reinterpret_cast<const double2*>(sites + i);
    const double2 a = __ldcs(src), b = __ldcs(src + 1), c = __ldcs(src + 2), d = __ldcs(src + 3), e = __ldcs(src + 4);
    const uint32_t parent = (uint32_t)(__double_as_longlong(e.y) & 0xffffffffLL);
    const uint32_t daughter = (uint32_t)((unsigned long long)__double_as_longlong(e.y) >> 32);
    out.x[pos] = a.x; out.y[pos] = a.y; out.z[pos] = b.x;
    out.ux[pos] = b.y; out.uy[pos] = c.x; out.uz[pos] = c.y;
    out.E[pos] = d.x; out.wgt[pos] = d.y;
    if (out.wgt2) out.wgt2[pos] = e.x;
    out.id_a[pos] = in.id_a[parent];
    out.id_b[pos] = did ? did[i] : daughter;
    out.id_c[pos] = in.id_b ? in.id_b[parent] : in.id_a[parent];
  }
}

// the exact cancelators' side table follows the sites to their rows (ABL_PARENT_FIELDS columns of `cap` entries: x, y, z, Esmp,
// ux, uy, uz, previous previous energy, previous energy, was_virtual)
__global__ void __launch_bounds__(256) place_parent_info_kernel(const double* __restrict__ site_parent, uint64_t n_sites,
                                                                const uint32_t* __restrict__ inv, double* __restrict__ out, uint64_t cap) {
  for (uint64_t pos = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; pos < n_sites; pos += (uint64_t)gridDim.x * blockDim.x) {
    const double2* src = reinterpret_cast<const double2*>(site_parent + (size_t)ABL_PARENT_FIELDS * inv[pos]);
#pragma unroll
    for (int q = 0; q < ABL_PARENT_FIELDS / 2; q++) {
      const double2 a = __ldcs(src + q);
      out[(2 * q) * cap + pos] = a.x;
      out[(2 * q + 1) * cap + pos] = a.y;
    }
  }
}

// dst row k = src row rows[k]; wgt from wgts when given (abl_bank_gather_device)
__global__ void __launch_bounds__(256) bank_gather_kernel(BankView src, const uint32_t* __restrict__ rows, const double* __restrict__ wgts,
                                                          uint64_t n, BankView dst) {
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t r = rows[k];
    dst.x[k] = src.x[r]; dst.y[k] = src.y[r]; dst.z[k] = src.z[r];
    dst.ux[k] = src.ux[r]; dst.uy[k] = src.uy[r]; dst.uz[k] = src.uz[r];
    dst.E[k] = src.E[r];
    dst.wgt[k] = wgts ? wgts[k] : src.wgt[r];
    if (dst.wgt2) dst.wgt2[k] = src.wgt2 ? src.wgt2[r] : 0.;
    dst.id_a[k] = src.id_a[r]; dst.id_b[k] = src.id_b[r]; dst.id_c[k] = src.id_c[r];
  }
}

// ---- weights ------------------------------------------------------------------------------------------------------
// out[0..3] += Npos, Nneg, Wpos, Wneg   (Wneg accumulated as a positive number, as the reference does)
__global__ void __launch_bounds__(256) weight_stats_kernel(const double* __restrict__ w, uint64_t n, double* out) {
  double np = 0., nn = 0., wp = 0., wn = 0.;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const double v = w[i];
    if (v > 0.) { wp += v; np += 1.; } else { wn -= v; nn += 1.; }
  }
  double vals[4] = {np, nn, wp, wn};
  __shared__ double sm[8][4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    double v = vals[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[wid][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.;
    for (int k = 0; k < 8; k++) v += sm[k][threadIdx.x];
    atomicAdd(&out[threadIdx.x], v);
  }
}

// Weighted moments of the bank about a point o: out[0..4] += sum w, sum w (x - ox), sum w (y - oy), sum w (z - oz), sum w |r - o|^2.
// PowerIterator::compute_pair_dist_sqrd (src/power_iterator.cpp:637-663) is the double sum over all pairs
// sum_ij w_i w_j |r_i - r_j|^2 / (2 W^2), 1e14 terms at the bench size; it equals sum_i w_i |r_i - c|^2 / W about the weighted
// centroid c = sum w r / W, which is two passes of this kernel.
__global__ void __launch_bounds__(256) bank_moments_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                           const double* __restrict__ z, const double* __restrict__ w, uint64_t n,
                                                           double ox, double oy, double oz, double* out) {
  double vals[5] = {0., 0., 0., 0., 0.};
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const double wi = w[i], dx = x[i] - ox, dy = y[i] - oy, dz = z[i] - oz;
    vals[0] += wi;
    vals[1] += wi * dx;
    vals[2] += wi * dy;
    vals[3] += wi * dz;
    vals[4] += wi * (dx * dx + dy * dy + dz * dz);
  }
  __shared__ double sm[8][5];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 5; q++) {
    double v = vals[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[wid][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double v = 0.;
    for (int k = 0; k < 8; k++) v += sm[k][threadIdx.x];
    atomicAdd(&out[threadIdx.x], v);
  }
}

// noise.cpp:438-456: sum of |w| over complex weights; w /= d
__global__ void __launch_bounds__(256) weight_magnitude_kernel(const double* __restrict__ w, const double* __restrict__ w2, uint64_t n, double* out) {
  double v = 0.;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    v += sqrt(w[i] * w[i] + w2[i] * w2[i]);
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __shared__ double sv[8];
  if ((threadIdx.x & 31) == 0) sv[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.;
    for (int k = 0; k < 8; k++) t += sv[k];
    atomicAdd(out, t);
  }
}
__global__ void __launch_bounds__(256) divide_weights_kernel(double* w, double* w2, uint64_t n, double d) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    w[i] = w[i] / d;
    w2[i] = w2[i] / d;
  }
}

__global__ void __launch_bounds__(256) scale_weights_kernel(double* w, uint64_t n, double f) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) w[i] *= f;
}

__global__ void __launch_bounds__(256) to_particles_kernel(BankView b, uint64_t first_id) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += (uint64_t)gridDim.x * blockDim.x) {
    b.id_b[i] = b.id_c[i];  // family id
    b.id_a[i] = first_id + i;
  }
}

// ---- entropy --------------------------------------------------------------------------------------------------------
// nshared = bins held in the block's dynamic shared memory (the whole mesh, or 0: meshes too large for it add to global
// memory directly).  The entropy mesh is a few hundred bins: 1e7 atomics on them in L2 took 2.5 ms, privatised 0.1 ms.
__global__ void __launch_bounds__(256) entropy_bin_kernel(DevMesh3 m, BankView b, double* bins, double* total, int nshared) {
  extern __shared__ double ent_bins[];
  for (int i = threadIdx.x; i < nshared; i += blockDim.x) ent_bins[i] = 0.;
  __syncthreads();
  double tw = 0.;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += (uint64_t)gridDim.x * blockDim.x) {
    const int nx = (int)floor((b.x[i] - m.lowx) / m.dx);
    const int ny = (int)floor((b.y[i] - m.lowy) / m.dy);
    const int nz = (int)floor((b.z[i] - m.lowz) / m.dz);
    if (nx >= 0 && nx < m.Nx && ny >= 0 && ny < m.Ny && nz >= 0 && nz < m.Nz) {
      const double w = b.wgt[i];
      tw += w;
      const size_t bin = (size_t)(m.Ny * m.Nz) * nx + (size_t)m.Nz * ny + nz;
      if (nshared) atomicAdd(&ent_bins[bin], w);
      else red_add(bins + bin, w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nshared; i += blockDim.x)
    if (ent_bins[i] != 0.) red_add(bins + i, ent_bins[i]);
  for (int o = 16; o > 0; o >>= 1) tw += __shfl_down_sync(0xffffffffu, tw, o);
  if ((threadIdx.x & 31) == 0 && tw != 0.) atomicAdd(total, tw);
}

// ---- source mesh tally ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) score_source_kernel(DevTally t, BankView b) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += (uint64_t)gridDim.x * blockDim.x) {
    const int ii = (int)floor((b.x[i] - t.lowx) / t.dx);  // division, not *dx_inv (source_mesh_tally.cpp:38-40)
    const int jj = (int)floor((b.y[i] - t.lowy) / t.dy);
    const int kk = (int)floor((b.z[i] - t.lowz) / t.dz);
    const int l = tally_energy_bin(t, b.E[i]);
    if (l == -1) continue;
    if (ii >= 0 && ii < t.Nx && jj >= 0 && jj < t.Ny && kk >= 0 && kk < t.Nz) {
      double scr = 1. / t.net_weight;
      if (t.quantity == ABL_Q_IMAG_SOURCE) scr *= (b.wgt2 ? b.wgt2[i] : 0.);
      else scr *= b.wgt[i];
      red_add(t.gen + tally_index(t, l, ii, jj, kk), scr);
    }
  }
}

// ---- approximate mesh cancellation ---------------------------------------------------------------------------------
// Dense per-bin accumulators, zeroed before each use.  Everything is a SUM, so the bins of several GPUs combine with one
// all-reduce: the weights of each sign are accumulated separately ("a particle with positive weight is in the bin" <=>
// its positive sum is > 0), members are counted.
struct CancelBins {
  double *sum_pos, *sum_neg;    // wgt:  sum of the positive / of the negative weights
  double *sum_pos2, *sum_neg2;  // wgt2
  uint32_t* count;              // members
};
__device__ __forceinline__ long long cancel_key(const DevMesh3& m, double x, double y, double z, double E) {
  const int i = (int)floor((x - m.lowx) / m.dx);
  const int j = (int)floor((y - m.lowy) / m.dy);
  const int k = (int)floor((z - m.lowz) / m.dz);
  int l = -1;
  if (m.eedges) {
    for (int e = 0; e < m.Ne; e++)
      if (ldt(&m.eedges[e]) <= E && E <= ldt(&m.eedges[e + 1])) { l = e; break; }
  } else {
    l = 0;
  }
  if (i < 0 || j < 0 || k < 0 || l < 0) return -1;
  if (i >= m.Nx || j >= m.Ny || k >= m.Nz || l >= m.Ne) return -1;
  return (long long)l + (long long)m.Ne * ((long long)k + (long long)m.Nz * ((long long)j + (long long)m.Ny * (long long)i));
}
__global__ void __launch_bounds__(256) cancel_accumulate_kernel(DevMesh3 m, BankView b, CancelBins cb) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += (uint64_t)gridDim.x * blockDim.x) {
    const long long key = cancel_key(m, b.x[i], b.y[i], b.z[i], b.E[i]);
    if (key < 0) continue;
    const double w = b.wgt[i], w2 = b.wgt2 ? b.wgt2[i] : 0.;
    atomicAdd(&cb.count[key], 1u);
    if (w > 0.) red_add(&cb.sum_pos[key], w);
    else if (w < 0.) red_add(&cb.sum_neg[key], w);
    if (w2 > 0.) red_add(&cb.sum_pos2[key], w2);
    else if (w2 < 0.) red_add(&cb.sum_neg2[key], w2);
  }
}
// approximate_mesh_cancelator.cpp:147-190: a bin with more than one member and both signs present gives every member
// the bin's mean weight (per weight component)
__global__ void __launch_bounds__(256) cancel_apply_kernel(DevMesh3 m, BankView b, CancelBins cb) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += (uint64_t)gridDim.x * blockDim.x) {
    const long long key = cancel_key(m, b.x[i], b.y[i], b.z[i], b.E[i]);
    if (key < 0) continue;
    const uint32_t members = cb.count[key];
    if (members > 1) {
      const double N = (double)members;
      const double sp = cb.sum_pos[key], sn = cb.sum_neg[key];
      if (sp > 0. && sn < 0.) b.wgt[i] = (sp + sn) / N;
      if (b.wgt2) {
        const double sp2 = cb.sum_pos2[key], sn2 = cb.sum_neg2[key];
        if (sp2 > 0. && sn2 < 0.) b.wgt2[i] = (sp2 + sn2) / N;
      }
    }
  }
}
// re-zero only the bins that were touched (the mesh is usually far larger than the bank)
__global__ void __launch_bounds__(256) cancel_reset_kernel(DevMesh3 m, BankView b, CancelBins cb) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += (uint64_t)gridDim.x * blockDim.x) {
    const long long key = cancel_key(m, b.x[i], b.y[i], b.z[i], b.E[i]);
    if (key < 0) continue;
    cb.count[key] = 0;
    cb.sum_pos[key] = 0.;
    cb.sum_neg[key] = 0.;
    cb.sum_pos2[key] = 0.;
    cb.sum_neg2[key] = 0.;
  }
}

// ---- initial source ---------------------------------------------------------------------------------------------------
// include/utils/direction.hpp:44-64
__device__ __forceinline__ V3 make_direction_mu_phi(double mu, double phi) {
  if (mu < -1.) mu = -1.; else if (mu > 1.) mu = 1.;
  if (phi < 0.) phi = 0.; else if (phi > 2 * ABL_PI) phi = 2 * ABL_PI;
  double sn, cs;
  det_sincos(phi, &sn, &cs);
  return make_direction(sqrt(1. - mu * mu) * cs, sqrt(1. - mu * mu) * sn, mu);
}

__global__ void __launch_bounds__(128) sample_source_kernel(const DevProblem P, BankView out, uint64_t first_id, int* error) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < out.n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t hid = first_id + i;
    uint64_t rng = pcg_advance(P.seed_state, P.stride * hid, P.jump);
    int si = 0;
    if (P.nsources >= 2) si = rng_discrete(rng, P.source_cp, P.nsources);
    const abl_source* S = P.sources + si;
    const int dkind = ldt(&S->direction_kind);
    V3 u;
    if (dkind == ABL_DIR_ISOTROPIC) {  // isotropic.cpp:28-36
      const double mu = 2. * rng_rand(rng) - 1.;
      const double phi = 2. * ABL_PI * rng_rand(rng);
      u = make_direction_mu_phi(mu, phi);
    } else {
      u = {ldt(&S->dir[0]), ldt(&S->dir[1]), ldt(&S->dir[2])};  // mono_directional.hpp:38: the stored direction, no draw
      if (dkind == ABL_DIR_CONE) {  // cone.cpp:34-42: mu in [cos(aperture), 1], phi in [0, 2 pi], about the axis
        const double ca = ldt(&S->cos_aperture);
        const double mu = (1. - ca) * rng_rand(rng) + ca;
        const double phi = 2. * ABL_PI * rng_rand(rng);
        u = rotate_direction(u, mu, phi);
      }
    }
    // Source::generate_particle draws the energy once, then again until it lies inside (min_energy, max_energy), at most 201
    // more times (source.cpp:48-58); a mono-energetic source draws nothing
    double E = ldt(&S->energy);
    bool bad = false;
    const int ekind = ldt(&S->energy_kind);
    if (ekind != ABL_EN_MONO) {
      const double ea = ldt(&S->en_a), eb = ldt(&S->en_b);
      const double emin = P.min_energy, emax = ldt(&P.ebounds[P.G]);
      auto sample_energy = [&]() {
        const double xi1 = rng_rand(rng), xi2 = rng_rand(rng), xi3 = rng_rand(rng);  // maxwellian.cpp:34-42
        double sn, c;
        det_sincos(ABL_PI * xi3 / 2., &sn, &c);
        const double w = -ea * (det_log(xi1) + det_log(xi2) * c * c);
        if (ekind == ABL_EN_MAXWELLIAN) return w;
        return w + 0.25 * ea * ea * eb + (2. * rng_rand(rng) - 1.) * sqrt(ea * ea * eb * w);  // watt.cpp:42-47
      };
      E = sample_energy();
      int e_count = 0;
      do {
        if (e_count > 200) { bad = true; break; }
        E = sample_energy();
        e_count++;
      } while (E <= emin || emax <= E);
    }
    const bool is_box = ldt(&S->is_box) != 0;
    const double lx = ldt(&S->low[0]), ly = ldt(&S->low[1]), lz = ldt(&S->low[2]);
    const double hx = ldt(&S->hi[0]), hy = ldt(&S->hi[1]), hz = ldt(&S->hi[2]);
    V3 r;
    Cursor c;
    c.err = 0;
    auto sample_pos = [&]() {
      if (is_box) {  // box.cpp:37-42
        r.x = (hx - lx) * rng_rand(rng) + lx;
        r.y = (hy - ly) * rng_rand(rng) + ly;
        r.z = (hz - lz) * rng_rand(rng) + lz;
      } else {
        r = {lx, ly, lz};
      }
      c.token = 0;
      cursor_restart(P, c, r, u);
    };
    sample_pos();
    int guard = 0;
    while (c.cell < 0) {  // rejection until inside the geometry (source.cpp:61-70)
      if (!is_box && ++guard > 1) { bad = true; break; }
      sample_pos();
    }
    if (!bad && ldt(&S->fissile_only)) {  // source.cpp:72-86
      int count = 0;
      while (c.cell < 0 || !ldt(&P.fissile[c.mat])) {
        if (count == 201) { bad = true; break; }
        sample_pos();
        count++;
      }
    }
    if (bad) atomicCAS(error, 0, ABL_ERR_INVALID);
    out.x[i] = r.x; out.y[i] = r.y; out.z[i] = r.z;
    out.ux[i] = u.x; out.uy[i] = u.y; out.uz[i] = u.z;
    out.E[i] = E;
    out.wgt[i] = 1.0;
    if (out.wgt2) out.wgt2[i] = 0.;
    out.id_a[i] = hid;
    out.id_b[i] = hid;
    out.id_c[i] = rng;
  }
}

// ---- tally statistics --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tally_record_kernel(double* __restrict__ gen, double* __restrict__ avg,
                                                           double* __restrict__ var, uint64_t n, double multiplier, double dg) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const double old_avg = avg[i];
    const double val = gen[i] * multiplier;
    const double a = old_avg + (val - old_avg) / dg;
    avg[i] = a;
    double v = var[i];
    v = v + (((val - old_avg) * (val - a) - (v)) / dg);
    var[i] = v;
  }
}
__global__ void __launch_bounds__(256) tally_std_kernel(const double* __restrict__ var, double* __restrict__ out, uint64_t n, double dg) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = sqrt(var[i] / dg);
}

__global__ void smem_base_probe_kernel(unsigned long long* out) {
  extern __shared__ __align__(16) unsigned char probe_raw[];
  if (threadIdx.x == 0) *out = (unsigned long long)(uintptr_t)(void*)probe_raw;  // generic address of the dynamic shared memory
}

// ---- probes --------------------------------------------------------------------------------------------------------------
__global__ void find_cells_kernel(const DevProblem P, uint64_t n, const double* r3, const double* u3, int32_t* cell, int32_t* mat) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    Cursor c;
    c.err = 0;
    c.token = 0;
    const V3 r{r3[3 * i], r3[3 * i + 1], r3[3 * i + 2]}, u{u3[3 * i], u3[3 * i + 1], u3[3 * i + 2]};
    cursor_restart(P, c, r, u);
    cell[i] = c.cell;
    mat[i] = c.mat;
  }
}
__global__ void surface_probe_kernel(const DevProblem P, int si, uint64_t n, const double* r3, const double* u3, const int32_t* on,
                                     int32_t* sign, double* dist, double* norm3) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const Surf s = load_surface(P, si);
    const V3 r{r3[3 * i], r3[3 * i + 1], r3[3 * i + 2]}, u{u3[3 * i], u3[3 * i + 1], u3[3 * i + 2]};
    sign[i] = surf_sign(s, r, u);
    dist[i] = surf_distance(s, r, u, on[i] != 0);
    const V3 nn = surf_norm(s, r);
    norm3[3 * i] = nn.x;
    norm3[3 * i + 1] = nn.y;
    norm3[3 * i + 2] = nn.z;
  }
}
__global__ void rng_probe_kernel(const DevProblem P, uint64_t history_id, int n, uint32_t* out_u32, double* out_rand) {
  if (threadIdx.x || blockIdx.x) return;
  uint64_t s = pcg_advance(P.seed_state, P.stride * history_id, P.jump);
  for (int i = 0; i < n; i++) out_u32[i] = pcg_next(s);
  s = pcg_advance(P.seed_state, P.stride * history_id, P.jump);
  for (int i = 0; i < n; i++) out_rand[i] = rng_rand(s);
}
__global__ void math_probe_kernel(int n, const double* x, double* lg, double* sn, double* cs) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    lg[i] = det_log(x[i]);
    det_sincos(x[i], &sn[i], &cs[i]);
  }
}

// pcg32 state of every history of a bank: seed(seed); advance(stride * history_id) (particle.hpp:188-193)
__global__ void __launch_bounds__(256) seed_streams_kernel(const DevProblem P, const uint64_t* __restrict__ history_id, uint64_t n,
                                                           uint64_t* __restrict__ state) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    state[i] = pcg_advance(P.seed_state, P.stride * history_id[i], P.jump);
}

// ---- BasicExactMGCancelator (src/basic_exact_mg_cancelator.cpp) ---------------------------------------------------------------
// The per-particle part, in parallel: the bin and the material a fission site falls in (add_particle :68-108, get_material
// :178-196), its point density f(r | r_parent, Esmp) = exp(-Esmp d) / d^2 (get_f :198-224) and the smallest f over the eight
// corners of its bin (get_min_f :226-262, the beta of mode `minimum`).  The bins themselves -- an unordered_map walked in
// insertion-dependent order, sums taken particle by particle -- are replayed on the host from these columns (api.cu).
__device__ __forceinline__ double exact_f(double x, double y, double z, double px, double py, double pz, double Esmp) {
  const double d = sqrt((x - px) * (x - px) + (y - py) * (y - py) + (z - pz) * (z - pz));
  return (1. / (d * d)) * det_exp(-Esmp * d);
}
__device__ __forceinline__ double min_ref(double a, double b) { return (b < a) ? b : a; }  // std::min
__global__ void __launch_bounds__(128) exact_prepare_kernel(const DevProblem P, const DevMesh3 m, BankView b, const double* __restrict__ parent,
                                                            uint64_t pcap, uint64_t n, int32_t* __restrict__ key, int32_t* __restrict__ mat,
                                                            double* __restrict__ f, double* __restrict__ fmin) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (uint64_t)gridDim.x * blockDim.x) {
    const double x = b.x[q], y = b.y[q], z = b.z[q];
    const int i = (int)floor((x - m.lowx) / m.dx), j = (int)floor((y - m.lowy) / m.dy), k = (int)floor((z - m.lowz) / m.dz);
    if (i < 0 || j < 0 || k < 0 || i >= m.Nx || j >= m.Ny || k >= m.Nz) {
      key[q] = -1;
      continue;
    }
    key[q] = k + m.Nz * (j + m.Ny * i);
    Cursor c;
    c.err = 0;
    c.token = 0;
    cursor_restart(P, c, V3{x, y, z}, V3{1., 0., 0.});
    mat[q] = c.cell < 0 ? -1 : c.mat;
    const double px = parent[q], py = parent[pcap + q], pz = parent[2 * pcap + q], Esmp = parent[3 * pcap + q];
    f[q] = exact_f(x, y, z, px, py, pz, Esmp);
    const double Xl = m.lowx + i * m.dx, Xh = Xl + m.dx, Yl = m.lowy + j * m.dy, Yh = Yl + m.dy, Zl = m.lowz + k * m.dz, Zh = Zl + m.dz;
    const double f1 = exact_f(Xh, Yh, Zh, px, py, pz, Esmp), f2 = exact_f(Xh, Yh, Zl, px, py, pz, Esmp), f3 = exact_f(Xh, Yl, Zh, px, py, pz, Esmp),
                 f4 = exact_f(Xh, Yl, Zl, px, py, pz, Esmp), f5 = exact_f(Xl, Yh, Zh, px, py, pz, Esmp), f6 = exact_f(Xl, Yh, Zl, px, py, pz, Esmp),
                 f7 = exact_f(Xl, Yl, Zh, px, py, pz, Esmp), f8 = exact_f(Xl, Yl, Zl, px, py, pz, Esmp);
    fmin[q] = min_ref(min_ref(min_ref(f1, f2), min_ref(f3, f4)), min_ref(min_ref(f5, f6), min_ref(f7, f8)));
  }
}

// settings::rng on the device: pcg32 on stream 2 (increment 5; settings.cpp:116-119)
struct GlobalStreamMath {
  static __device__ __forceinline__ uint32_t next(uint64_t& state) {
    const uint64_t old = state;
    state = old * ABL_PCG_MULT + 5ULL;
    const uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    const uint32_t rot = (uint32_t)(old >> 59u);
    return __funnelshift_r(xorshifted, xorshifted, rot);
  }
  static __device__ __forceinline__ double rand(uint64_t& s) {
    const uint32_t lo = next(s);
    const uint32_t hi = next(s);
    double sum = (double)lo;
    sum += (double)hi * 4294967296.0;
    double ret = sum / 18446744073709551616.0;
    if (ret >= 1.0) ret = 0.99999999999999988897769753748;
    return ret;
  }
  static __device__ __forceinline__ double div(double a, double b) { return a / b; }
  static __device__ __forceinline__ double sqrt_(double a) { return sqrt(a); }
};

// get_averages / get_averages_sobol (:264-364) for the bins that need them, one thread per bin: n_samples points of the bin that
// lie in the bin's material -- from the Sobol sequence (restarted in every bin) or from the global engine advanced to the bin's
// offset --, then for every particle of the bin the mean of f and of 1/f over those points.  desc rows: key, material, first row
// in `rows`, count, engine offset.  A bin that cannot place a point within 100 tries per sample is marked (can_cancel = false).
#define ABL_EXACT_MAX_SAMPLES 64
struct SobolMatrices3 {  // (four dimensions: the `type: exact` cancelator takes the energy group of a point from the fourth)
  unsigned long long m[4][52];
};
__device__ __forceinline__ double sobol_sample(const SobolMatrices3& M, unsigned long long index, int dim) {
  unsigned long long result = 0;
  for (int i = 0; index; index >>= 1, ++i)
    if (index & 1) result ^= M.m[dim][i];
  return (double)result * (1.0 / 4503599627370496.0);  // 2^-52
}
__device__ __forceinline__ uint64_t stream_advance(uint64_t state, uint64_t delta) {  // pcg advance with increment 5
  uint64_t acc_mult = 1, acc_plus = 0, cur_mult = ABL_PCG_MULT, cur_plus = 5ULL;
  while (delta > 0) {
    if (delta & 1) {
      acc_mult *= cur_mult;
      acc_plus = acc_plus * cur_mult + cur_plus;
    }
    cur_plus = (cur_mult + 1) * cur_plus;
    cur_mult *= cur_mult;
    delta >>= 1;
  }
  return acc_mult * state + acc_plus;
}
__global__ void __launch_bounds__(64) exact_average_kernel(const DevProblem P, const DevMesh3 m, const SobolMatrices3 SM, uint64_t nbins,
                                                           const unsigned long long* __restrict__ desc, const unsigned long long* __restrict__ rows,
                                                           BankView b, const double* __restrict__ parent, uint64_t pcap, int nsamples, int use_sobol,
                                                           uint64_t rng0, double* __restrict__ avg_f, double* __restrict__ avg_finv,
                                                           int32_t* __restrict__ can_cancel) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nbins; q += (uint64_t)gridDim.x * blockDim.x) {
    const int key = (int)desc[5 * q], mat = (int)(long long)desc[5 * q + 1];
    const uint64_t first = desc[5 * q + 2], count = desc[5 * q + 3];
    uint64_t rng = stream_advance(rng0, desc[5 * q + 4]);
    const int k = key % m.Nz, j = (key / m.Nz) % m.Ny, i = key / (m.Nz * m.Ny);
    const double Xl = m.lowx + i * m.dx, Yl = m.lowy + j * m.dy, Zl = m.lowz + k * m.dz;
    double rs[ABL_EXACT_MAX_SAMPLES][3];
    unsigned long long sobol_index = 0;
    bool ok_all = true;
    for (int sidx = 0; sidx < nsamples && ok_all; sidx++) {
      bool ok = false;
      for (int t = 0; t < 100 && !ok; t++) {
        double x, y, z;
        if (use_sobol) {
          x = Xl + sobol_sample(SM, sobol_index, 0) * m.dx;
          y = Yl + sobol_sample(SM, sobol_index, 1) * m.dy;
          z = Zl + sobol_sample(SM, sobol_index, 2) * m.dz;
          sobol_index++;
        } else {
          x = Xl + GlobalStreamMath::rand(rng) * m.dx;
          y = Yl + GlobalStreamMath::rand(rng) * m.dy;
          z = Zl + GlobalStreamMath::rand(rng) * m.dz;
        }
        Cursor c;
        c.err = 0;
        c.token = 0;
        cursor_restart(P, c, V3{x, y, z}, V3{1., 0., 0.});
        ok = (c.cell < 0 ? -1 : c.mat) == mat;
        rs[sidx][0] = x; rs[sidx][1] = y; rs[sidx][2] = z;
      }
      ok_all = ok;
    }
    can_cancel[q] = ok_all ? 1 : 0;
    if (!ok_all) continue;
    for (uint64_t e = 0; e < count; e++) {
      const uint64_t row = rows[first + e];
      const double px = parent[row], py = parent[pcap + row], pz = parent[2 * pcap + row], Esmp = parent[3 * pcap + row];
      double sum_f = 0., sum_f_inv = 0.;
      for (int sidx = 0; sidx < nsamples; sidx++) {
        const double f = exact_f(rs[sidx][0], rs[sidx][1], rs[sidx][2], px, py, pz, Esmp);
        sum_f += f;
        sum_f_inv += 1. / f;
      }
      avg_f[first + e] = sum_f / (double)nsamples;
      avg_finv[first + e] = sum_f_inv / (double)nsamples;
    }
  }
}

// ---- ExactMGCancelator (src/exact_mg_cancelator.cpp; cancelator: {type: exact}) ------------------------------------------------
// The same division of labour as for basic-exact: per-particle densities and per-bin averages here, the unordered_map of bins
// replayed on the host.  The density of a site also carries the parent's scattering-angle pdf towards the site and, with a chi
// matrix, the chi element (get_f :215-235); bins are keyed by mesh cell and energy bin (get_key :115-152).
__device__ __forceinline__ int group_closed(const DevProblem& P, double E) {  // settings::group (src/settings.cpp:95-104)
  for (int g = 0; g < P.G; g++)
    if (P.ebounds[g] <= E && E <= P.ebounds[g + 1]) return g;
  return 0;
}
__device__ inline double angle_pdf_at(const DevProblem& P, int mat, int g1, int g3, double x) {  // MGAngleDistribution::pdf (:62-75)
  const abl_angle_table at = P.angle[((size_t)mat * P.G + g1) * P.G + g3];
  if (at.n < 0) {  // the default isotropic table {-1, 1} / {0.5, 0.5}
    if (x < -1. || x > 1. || x == -1. || x == 1.) return 0.5;
    return 0. * (x - -1.) + 0.5;
  }
  const double* mu = P.amu + at.offset;
  const double* pdf = P.apdf + at.offset;
  if (x < mu[0]) return pdf[0];
  if (x > mu[at.n - 1]) return pdf[at.n - 1];
  int lo = 0, len = at.n;
  while (len > 0) {  // std::lower_bound
    const int half = len >> 1;
    if (mu[lo + half] < x) {
      lo = lo + half + 1;
      len = len - half - 1;
    } else {
      len = half;
    }
  }
  if (x == mu[lo]) return pdf[lo];
  lo--;
  const double m = (pdf[lo + 1] - pdf[lo]) / (mu[lo + 1] - mu[lo]);
  return m * (x - mu[lo]) + pdf[lo];
}
__device__ inline double exact_full_f(const DevProblem& P, int mat, double r1x, double r1y, double r1z, double u1x, double u1y, double u1z,
                                      int g1, int g3, double r4x, double r4y, double r4z, int g4, double Esmp) {
  const double dx = r4x - r1x, dy = r4y - r1y, dz = r4z - r1z;
  const double d = sqrt(dx * dx + dy * dy + dz * dz);
  const V3 u = make_direction(dx, dy, dz);
  const double mu = u.x * u1x + u.y * u1y + u.z * u1z;
  const double pdf_mu = angle_pdf_at(P, mat, g1, g3, mu);
  const double pdf_chi = P.chi_matrix ? P.chi_pdf[((size_t)mat * P.G + g3) * P.G + g4] : 1.;
  return (pdf_mu * pdf_chi / (d * d)) * det_exp(-Esmp * d);
}
// energy bin of group g: the first bin that lists it; the number of bins when none does (get_key :133-147)
__device__ inline int exact_energy_bin(const int32_t* egb, int g, bool& found) {
  const int nbins = egb ? egb[0] : 0;
  int pos = 1;
  found = false;
  for (int e = 0; e < nbins; e++) {
    const int cnt = egb[pos];
    for (int q = 0; q < cnt; q++)
      if (egb[pos + 1 + q] == g) { found = true; return e; }
    pos += 1 + cnt;
  }
  return nbins;
}
__device__ inline const int32_t* exact_bin_groups(const int32_t* egb, int e, int& cnt) {
  int pos = 1;
  for (int q = 0; q < e; q++) pos += 1 + egb[pos];
  cnt = egb[pos];
  return egb + pos + 1;
}
__global__ void __launch_bounds__(128) exact_full_prepare_kernel(const DevProblem P, const DevMesh3 m, BankView b, const double* __restrict__ parent,
                                                                 uint64_t pcap, uint64_t n, int32_t* __restrict__ key, int32_t* __restrict__ mat,
                                                                 double* __restrict__ f) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (uint64_t)gridDim.x * blockDim.x) {
    const double x = b.x[q], y = b.y[q], z = b.z[q];
    const int g4 = group_closed(P, b.E[q]);
    key[q] = -1;
    if (x < m.lowx || x > m.hix || y < m.lowy || y > m.hiy || z < m.lowz || z > m.hiz) continue;
    const long long i = (long long)floor((x - m.lowx) / m.dx), j = (long long)floor((y - m.lowy) / m.dy), k = (long long)floor((z - m.lowz) / m.dz);
    bool found;
    const int e = exact_energy_bin(P.egb, g4, found);
    if (!found && P.chi_matrix) continue;
    key[q] = (int32_t)(e + (long long)m.Ne * (k + (long long)m.Nz * (j + (long long)m.Ny * i)));
    Cursor c;
    c.err = 0;
    c.token = 0;
    cursor_restart(P, c, V3{x, y, z}, V3{1., 0., 0.});
    const int mt = c.cell < 0 ? -1 : c.mat;
    mat[q] = mt;
    if (mt < 0) { f[q] = 0.; continue; }
    const int g1 = group_closed(P, parent[7 * pcap + q]), g3 = group_closed(P, parent[8 * pcap + q]);
    f[q] = exact_full_f(P, mt, parent[q], parent[pcap + q], parent[2 * pcap + q], parent[4 * pcap + q], parent[5 * pcap + q], parent[6 * pcap + q],
                        g1, g3, x, y, z, g4, parent[3 * pcap + q]);
  }
}
// compute_averages (:296-364): one thread per bin that holds both signs.  desc rows: key, material, first row in `rows`, count, unused.
__global__ void __launch_bounds__(64) exact_full_average_kernel(const DevProblem P, const DevMesh3 m, const SobolMatrices3 SM, uint64_t nbins,
                                                                const unsigned long long* __restrict__ desc, const unsigned long long* __restrict__ rows,
                                                                const double* __restrict__ parent, uint64_t pcap, int nsamples,
                                                                double* __restrict__ avg_f, double* __restrict__ avg_finv, int32_t* __restrict__ can_cancel) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nbins; q += (uint64_t)gridDim.x * blockDim.x) {
    const long long hk = (long long)desc[5 * q];
    const int mat = (int)(long long)desc[5 * q + 1];
    const uint64_t first = desc[5 * q + 2], count = desc[5 * q + 3];
    const int e = (int)(hk % m.Ne);
    const long long k = (hk / m.Ne) % m.Nz, j = (hk / ((long long)m.Ne * m.Nz)) % m.Ny, i = hk / ((long long)m.Ne * m.Nz * m.Ny);
    const double Xl = m.lowx + (double)i * m.dx, Yl = m.lowy + (double)j * m.dy, Zl = m.lowz + (double)k * m.dz;
    double rs[ABL_EXACT_MAX_SAMPLES][3];
    int gs[ABL_EXACT_MAX_SAMPLES];
    unsigned long long sobol_index = 0;
    bool ok_all = mat >= 0;
    for (int sidx = 0; sidx < nsamples && ok_all; sidx++) {  // sample_point :249-294
      bool ok = false;
      for (int t = 0; t < 100 && !ok; t++) {
        const double x = Xl + sobol_sample(SM, sobol_index, 0) * m.dx;
        const double y = Yl + sobol_sample(SM, sobol_index, 1) * m.dy;
        const double z = Zl + sobol_sample(SM, sobol_index, 2) * m.dz;
        sobol_index++;
        Cursor c;
        c.err = 0;
        c.token = 0;
        cursor_restart(P, c, V3{x, y, z}, V3{1., 0., 0.});
        ok = (c.cell < 0 ? -1 : c.mat) == mat;
        rs[sidx][0] = x; rs[sidx][1] = y; rs[sidx][2] = z;
      }
      gs[sidx] = 0;
      if (ok && P.chi_matrix) {
        const double xi = sobol_sample(SM, sobol_index - 1, 3);
        int cnt;
        const int32_t* groups = exact_bin_groups(P.egb, e, cnt);
        gs[sidx] = groups[(int)floor(xi * (double)cnt)];
      }
      ok_all = ok;
    }
    for (uint64_t pe = 0; pe < count && ok_all; pe++) {
      const uint64_t row = rows[first + pe];
      const int g1 = group_closed(P, parent[7 * pcap + row]), g3 = group_closed(P, parent[8 * pcap + row]);
      double sum_f = 0., sum_f_inv = 0.;
      for (int sidx = 0; sidx < nsamples; sidx++) {
        const double f = exact_full_f(P, mat, parent[row], parent[pcap + row], parent[2 * pcap + row], parent[4 * pcap + row], parent[5 * pcap + row],
                                      parent[6 * pcap + row], g1, g3, rs[sidx][0], rs[sidx][1], rs[sidx][2], gs[sidx], parent[3 * pcap + row]);
        if (f == 0.) {  // "bin.can_cancel = false; return;"
          ok_all = false;
          break;
        }
        sum_f += f;
        sum_f_inv += 1. / f;
      }
      avg_f[first + pe] = sum_f / (double)nsamples;
      avg_finv[first + pe] = sum_f_inv / (double)nsamples;
    }
    can_cancel[q] = ok_all ? 1 : 0;
  }
}
// get_new_particles of both exact cancelators on one warp.  The uniform particles are drawn one after the other from the one global
// engine, and how many draws a particle takes depends on how many position tries it needs -- but almost every particle needs one
// (its bin holds its material almost everywhere).  So lane l assumes that the l particles before it took one try each, jumps the
// engine to its own offset and draws its particle; the lanes up to the first one whose first try missed are right and are
// committed, that lane finishes its particle serially, and the next round starts behind it.  The bank is the reference's,
// particle for particle.  cum: exclusive prefix sums of the list's N column (uniform particles before every entry).
template <bool FULL>
__global__ void __launch_bounds__(32) exact_uniform_warp_kernel(const DevProblem P, const DevMesh3 m, const double* __restrict__ list,
                                                                const unsigned long long* __restrict__ cum, uint64_t nlist, uint64_t total,
                                                                uint64_t* rng_state, BankView b, uint64_t first, uint64_t capacity,
                                                                double* parent, uint64_t pcap, unsigned long long* out2) {
  const int lane = threadIdx.x;
  // engine outputs one particle consumes when its first position try hits: 3 position draws, the energy draw (if any), the two
  // direction draws and, for basic-exact (MGNuclide::sample_fission), the delayed-neutron draw; two outputs per draw
  const int energy_draw = FULL ? ((P.chi_matrix || P.G >= 2) ? 1 : 0) : (P.G >= 2 ? 1 : 0);
  const uint64_t D2 = 2ULL * (uint64_t)(3 + energy_draw + 2 + (FULL ? 0 : 1));
  uint64_t rng_base = *rng_state;
  uint64_t base_q = 0;
  unsigned long long failed = 0;
  // the engine's jump over lane * D2 outputs (state' = jm * state + jp), once: the O(log) advance costs more than a round otherwise
  uint64_t jm = 1, jp = 0;
  {
    uint64_t cur_mult = ABL_PCG_MULT, cur_plus = 5ULL, delta = (uint64_t)lane * D2;
    while (delta > 0) {
      if (delta & 1) {
        jm *= cur_mult;
        jp = jp * cur_mult + cur_plus;
      }
      cur_plus = (cur_mult + 1) * cur_plus;
      cur_mult *= cur_mult;
      delta >>= 1;
    }
  }
  while (base_q < total && !failed) {
    const uint64_t q = base_q + (uint64_t)lane;
    const bool active = q < total;
    uint64_t rng = jm * rng_base + jp;
    bool ok = false;
    V3 r{0., 0., 0.};
    int mat = -1, e = 0;
    double w = 0., w2 = 0., Xl = 0., Yl = 0., Zl = 0.;
    if (active) {
      uint64_t lo = 0, hi = nlist;  // the entry with cum[entry] <= q < cum[entry + 1]
      while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (cum[mid] <= q) lo = mid; else hi = mid;
      }
      const long long hk = (long long)list[5 * lo];
      mat = (int)list[5 * lo + 1];
      w = list[5 * lo + 3];
      w2 = list[5 * lo + 4];
      long long i, j, k;
      if (FULL) {
        e = (int)(hk % m.Ne);
        k = (hk / m.Ne) % m.Nz; j = (hk / ((long long)m.Ne * m.Nz)) % m.Ny; i = hk / ((long long)m.Ne * m.Nz * m.Ny);
      } else {
        k = hk % m.Nz; j = (hk / m.Nz) % m.Ny; i = hk / ((long long)m.Nz * m.Ny);
      }
      Xl = m.lowx + (double)i * m.dx; Yl = m.lowy + (double)j * m.dy; Zl = m.lowz + (double)k * m.dz;
      const double x = Xl + GlobalStreamMath::rand(rng) * m.dx;
      const double y = Yl + GlobalStreamMath::rand(rng) * m.dy;
      const double z = Zl + GlobalStreamMath::rand(rng) * m.dz;
      r = V3{x, y, z};
      Cursor c;
      c.err = 0;
      c.token = 0;
      cursor_restart(P, c, r, V3{1., 0., 0.});
      ok = (c.cell < 0 ? -1 : c.mat) == mat;
    }
    const unsigned miss = __ballot_sync(0xffffffffu, active && !ok);
    const unsigned act = __ballot_sync(0xffffffffu, active);
    const int first_miss = miss ? __ffs(miss) - 1 : 32;
    const int nact = __popc(act);
    const bool mine = active && lane <= first_miss;  // lanes before the first miss are right; that lane goes on alone
    if (mine && lane == first_miss) {
      for (int t = 1; t < 100 && !ok; t++) {
        const double x = Xl + GlobalStreamMath::rand(rng) * m.dx;
        const double y = Yl + GlobalStreamMath::rand(rng) * m.dy;
        const double z = Zl + GlobalStreamMath::rand(rng) * m.dz;
        r = V3{x, y, z};
        Cursor c;
        c.err = 0;
        c.token = 0;
        cursor_restart(P, c, r, V3{1., 0., 0.});
        ok = (c.cell < 0 ? -1 : c.mat) == mat;
      }
      if (!ok) failed = 1;  // "Couldn't sample position for uniform particle."
    }
    if (mine && ok) {
      int e_index = 0;
      V3 dir;
      if (FULL) {  // ExactMGCancelator::get_new_particles (:511-590)
        if (P.chi_matrix) {
          const double xi_E = GlobalStreamMath::rand(rng);
          int cnt;
          const int32_t* groups = exact_bin_groups(P.egb, e, cnt);
          e_index = groups[(int)floor(xi_E * (double)cnt)];
        } else if (P.G >= 2) {
          e_index = rng_discrete<GlobalStreamMath>(rng, P.chi_cp + (size_t)mat * P.G * P.G, P.G);
        }
        // Direction u_smp(2 rand - 1, 2 pi rand): g++ evaluates the two arguments from the right, so phi takes the first draw
        double phi = 2. * ABL_PI * GlobalStreamMath::rand(rng);
        double mu = 2. * GlobalStreamMath::rand(rng) - 1.;
        if (mu < -1.) mu = -1.; else if (mu > 1.) mu = 1.;
        if (phi < 0.) phi = 0.; else if (phi > 2 * ABL_PI) phi = 2 * ABL_PI;
        double sn, cs;
        det_sincos(phi, &sn, &cs);
        dir = make_direction(sqrt(1. - mu * mu) * cs, sqrt(1. - mu * mu) * sn, mu);
      } else {  // BasicExactMGCancelator::get_new_particles (:557-617): MGNuclide::sample_fission(0, +z, group 0, P_delayed 0)
        if (P.G >= 2) e_index = rng_discrete<GlobalStreamMath>(rng, P.chi_cp + (size_t)mat * P.G * P.G, P.G);
        const double mu = 2. * GlobalStreamMath::rand(rng) - 1.;
        const double phi = 2. * ABL_PI * GlobalStreamMath::rand(rng);
        dir = rotate_dir<GlobalStreamMath>(V3{0., 0., 1.}, mu, phi);
        (void)GlobalStreamMath::rand(rng);  // the delayed-neutron draw against P_delayed = 0
      }
      const uint64_t row = first + q;
      if (row < capacity) {
        b.x[row] = r.x; b.y[row] = r.y; b.z[row] = r.z;
        b.ux[row] = dir.x; b.uy[row] = dir.y; b.uz[row] = dir.z;
        b.E[row] = ldt(&P.gmid[e_index]);
        b.wgt[row] = w;
        if (b.wgt2) b.wgt2[row] = w2;
        b.id_a[row] = 0; b.id_b[row] = 0; b.id_c[row] = 0;
        if (row < pcap)
          for (int f = 0; f < ABL_PARENT_FIELDS; f++) parent[f * pcap + row] = f == 4 ? 1. : 0.;
      }
    }
    failed = __any_sync(0xffffffffu, failed != 0) ? 1 : 0;
    if (first_miss < 32) {  // the engine continues where the lane that went on alone left it
      rng_base = __shfl_sync(0xffffffffu, rng, first_miss);
      base_q += (uint64_t)first_miss + 1;
    } else {
      rng_base = stream_advance(rng_base, (uint64_t)nact * D2);
      base_q += (uint64_t)nact;
    }
  }
  if (lane == 0) {
    *rng_state = rng_base;
    out2[0] = first + total;
    out2[1] = failed;
  }
}

}  // namespace abl
