// A1 on the device — candidate vacancy-hop enumeration for many replicas in one launch.
//
// Restates rlsim/actions/action.py:12-61 (`get_action_space`) for orthorhombic cells, generalised to k vacancies exactly
// like the host mirror rlvacdiffsim_b200/actions.py: atoms are mapped to the perfect FCC sites of
// bulk(..., cubic=True).repeat(round(diag / a)) by minimum-image argmin (action.py:27-30), the unoccupied sites are the
// vacancies (:33-37), and every occupied site within (0.1, 1.1 a/sqrt2) of a vacancy yields the row
// [atom_idx, minimum-image displacement to the vacant site] (:41-59), vacancies and sites in ascending index order.
// All arithmetic is fp64 with the host's operation order (d/L, subtract rint, times L; sqrt((x^2+y^2)+z^2)), every
// multiply-add left uncontracted, so indices and order are bit-identical and displacements agree to the last ulp of
// numpy's solve.  One CTA per replica; sites are generated on the fly; no atomics except the "last atom wins" site
// claim (atomicMax), which is what the reference's dict comprehension does.
#include "common.cuh"

namespace rlvd {

namespace {

constexpr int AT = 256;            // threads
constexpr int MAX_SITES = 8192;    // 32 KB of shared site -> atom map (up to 12 x 12 x 12 conventional cells... 4*n^3)

__device__ __forceinline__ void site_pos(int s, int ny, int nz, double a, double& x, double& y, double& z) {
    // site s = ((i * ny + j) * nz + k) * 4 + b, basis (0,0,0) (0,.5,.5) (.5,0,.5) (.5,.5,0)   (structures.fcc_sites)
    const int b = s & 3, c = s >> 2;
    const int k = c % nz, j = (c / nz) % ny, i = c / (nz * ny);
    const double bx = (b == 2 || b == 3) ? 0.5 : 0.0, by = (b == 1 || b == 3) ? 0.5 : 0.0, bz = (b == 1 || b == 2) ? 0.5 : 0.0;
    x = __dmul_rn((double)i + bx, a);
    y = __dmul_rn((double)j + by, a);
    z = __dmul_rn((double)k + bz, a);
}

// minimum image of d along one orthorhombic axis, numpy order: frac = d / L; frac -= rint(frac); frac * L
__device__ __forceinline__ double mic1(double d, double L) {
    double f = __ddiv_rn(d, L);
    f = __dsub_rn(f, rint(f));
    return __dmul_rn(f, L);
}
__device__ __forceinline__ double norm3(double x, double y, double z) {
    return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}

__global__ void __launch_bounds__(AT) action_space_kernel(int n_struct, const int* __restrict__ node_ptr,
                                                          const double* __restrict__ pos, const double* __restrict__ cell,
                                                          double a, int max_vac, int max_actions, int* __restrict__ count,
                                                          double* __restrict__ actions, int* __restrict__ status) {
    __shared__ int site_atom[MAX_SITES];
    __shared__ int vac[64];
    __shared__ int n_vac_s, n_act_s;
    __shared__ int part[AT + 1];
    const int g = blockIdx.x, tid = threadIdx.x;
    if (g >= n_struct) return;
    const int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
    const double* C = cell + 9 * (int64_t)g;
    const double Lx = C[0], Ly = C[4], Lz = C[8];
    const bool diag = C[1] == 0.0 && C[2] == 0.0 && C[3] == 0.0 && C[5] == 0.0 && C[6] == 0.0 && C[7] == 0.0;
    const int nx = (int)rint(Lx / a), ny = (int)rint(Ly / a), nz = (int)rint(Lz / a);
    const int ns = 4 * nx * ny * nz;
    if (tid == 0) { count[g] = 0; n_vac_s = 0; n_act_s = 0; }
    if (!diag || ns <= 0 || ns > MAX_SITES) {
        if (tid == 0) status[g] = 1;                       // unsupported cell (host enumerator handles it)
        return;
    }
    for (int s = tid; s < ns; s += AT) site_atom[s] = -1;
    __syncthreads();
    // ---- action.py:27-30: atom -> nearest site (first minimum, like np.argmin); dict semantics: last atom wins
    for (int i = tid; i < n; i += AT) {
        const double px = pos[3 * (int64_t)(n0 + i)], py = pos[3 * (int64_t)(n0 + i) + 1], pz = pos[3 * (int64_t)(n0 + i) + 2];
        double best = 1e300;
        int bs = 0;
        for (int s = 0; s < ns; ++s) {
            double sx, sy, sz;
            site_pos(s, ny, nz, a, sx, sy, sz);
            const double d = norm3(mic1(__dsub_rn(sx, px), Lx), mic1(__dsub_rn(sy, py), Ly), mic1(__dsub_rn(sz, pz), Lz));
            if (d < best) { best = d; bs = s; }
        }
        atomicMax(&site_atom[bs], i);
    }
    __syncthreads();
    // ---- action.py:33-37: vacant sites in ascending order
    if (tid == 0) {
        int nv = 0;
        for (int s = 0; s < ns; ++s)
            if (site_atom[s] < 0) {
                if (nv < 64) vac[nv] = s;
                ++nv;
            }
        n_vac_s = nv;
    }
    __syncthreads();
    const int nv = n_vac_s;
    if (nv < 1 || nv > max_vac || nv > 64) {
        if (tid == 0) status[g] = 2;                       // action.py:35 assert (generalised to 1..max_vacancies)
        return;
    }
    // ---- action.py:41-59: occupied nearest-neighbour sites of every vacancy, ascending site order
    const double nn_hi = __dmul_rn(__ddiv_rn(a, sqrt(2.0)), 1.1);
    const int chunk = (ns + AT - 1) / AT;
    for (int v = 0; v < nv; ++v) {
        double vx, vy, vz;
        site_pos(vac[v], ny, nz, a, vx, vy, vz);
        const int s_lo = tid * chunk, s_hi = min(ns, s_lo + chunk);
        int mine = 0;
        for (int s = s_lo; s < s_hi; ++s) {
            double sx, sy, sz;
            site_pos(s, ny, nz, a, sx, sy, sz);
            const double d = norm3(mic1(__dsub_rn(sx, vx), Lx), mic1(__dsub_rn(sy, vy), Ly), mic1(__dsub_rn(sz, vz), Lz));
            if (d > 0.1 && d < nn_hi && site_atom[s] >= 0) ++mine;
        }
        part[tid] = mine;
        __syncthreads();
        if (tid == 0) {                                    // tiny exclusive scan; keeps the output order deterministic
            int run = n_act_s;
            for (int t = 0; t < AT; ++t) { const int c = part[t]; part[t] = run; run += c; }
            part[AT] = run;
        }
        __syncthreads();
        int w = part[tid];
        for (int s = s_lo; s < s_hi; ++s) {
            double sx, sy, sz;
            site_pos(s, ny, nz, a, sx, sy, sz);
            const double d = norm3(mic1(__dsub_rn(sx, vx), Lx), mic1(__dsub_rn(sy, vy), Ly), mic1(__dsub_rn(sz, vz), Lz));
            if (d > 0.1 && d < nn_hi && site_atom[s] >= 0) {
                const int at = site_atom[s];
                if (w < max_actions) {
                    double* o = actions + ((int64_t)g * max_actions + w) * 4;
                    o[0] = (double)at;
                    o[1] = mic1(__dsub_rn(vx, pos[3 * (int64_t)(n0 + at)]), Lx);          // action.py:53-57
                    o[2] = mic1(__dsub_rn(vy, pos[3 * (int64_t)(n0 + at) + 1]), Ly);
                    o[3] = mic1(__dsub_rn(vz, pos[3 * (int64_t)(n0 + at) + 2]), Lz);
                }
                ++w;
            }
        }
        __syncthreads();
        if (tid == 0) n_act_s = part[AT];
        __syncthreads();
    }
    if (tid == 0) {
        count[g] = min(n_act_s, max_actions);
        status[g] = n_act_s > max_actions ? 3 : 0;         // 3: more candidates than the caller's buffer holds
    }
}


// ---- general (non-orthorhombic) cells: the demo POSCAR_0 of the reference is rhombohedral -----------------------------
// Same steps as above with the host mirror's general minimum image (rlvacdiffsim_b200/actions.py `mic_vectors`, which
// restates ase.geometry.find_mic's result): fractional rounding, then the 27 images around it, a candidate replacing the
// incumbent only if it is shorter by more than 1e-12.  The perfect-lattice sites of a cell are a property of the CELL, not
// of the state: the caller computes them once on the host (`lattice_sites_for_cell`: conventional FCC sites clipped to the
// cell, reference order) and passes them in.  The final displacement is the reference's formula (action.py:53-57:
// frac = solve(cell.T, diff); frac -= round(frac); cell.T @ frac) evaluated with the fp64 inverse cell.
__device__ __forceinline__ double mic_general(const double* C, const double* Ci, double dx, double dy, double dz) {
    double f0 = dx * Ci[0] + dy * Ci[3] + dz * Ci[6], f1 = dx * Ci[1] + dy * Ci[4] + dz * Ci[7], f2 = dx * Ci[2] + dy * Ci[5] + dz * Ci[8];
    f0 -= rint(f0); f1 -= rint(f1); f2 -= rint(f2);
    const double bx = f0 * C[0] + f1 * C[3] + f2 * C[6], by = f0 * C[1] + f1 * C[4] + f2 * C[7], bz = f0 * C[2] + f1 * C[5] + f2 * C[8];
    double best = bx * bx + by * by + bz * bz;
    for (int s0 = -1; s0 <= 1; ++s0)
        for (int s1 = -1; s1 <= 1; ++s1)
            for (int s2 = -1; s2 <= 1; ++s2) {
                if ((s0 | s1 | s2) == 0) continue;
                const double cx = bx + s0 * C[0] + s1 * C[3] + s2 * C[6], cy = by + s0 * C[1] + s1 * C[4] + s2 * C[7],
                             cz = bz + s0 * C[2] + s1 * C[5] + s2 * C[8];
                const double n2 = cx * cx + cy * cy + cz * cz;
                if (n2 < best - 1e-12) best = n2;
            }
    return sqrt(best);
}

__global__ void __launch_bounds__(AT) action_space_general_kernel(int n_struct, const int* __restrict__ node_ptr, const double* __restrict__ pos,
                                                                  const double* __restrict__ cell, const double* __restrict__ inv_cell,
                                                                  const int* __restrict__ site_ptr, const double* __restrict__ sites,
                                                                  double a, int max_vac, int max_actions, int* __restrict__ count,
                                                                  double* __restrict__ actions, int* __restrict__ status) {
    __shared__ int site_atom[MAX_SITES];
    __shared__ int vac[64];
    __shared__ int n_vac_s, n_act_s;
    __shared__ int part[AT + 1];
    __shared__ double Cs[9], Cis[9];
    const int g = blockIdx.x, tid = threadIdx.x;
    if (g >= n_struct) return;
    const int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
    const int s0 = site_ptr[g], ns = site_ptr[g + 1] - s0;
    if (tid < 9) { Cs[tid] = cell[9 * (int64_t)g + tid]; Cis[tid] = inv_cell[9 * (int64_t)g + tid]; }
    if (tid == 0) { count[g] = 0; n_vac_s = 0; n_act_s = 0; }
    if (ns <= 0 || ns > MAX_SITES) {
        if (tid == 0) status[g] = 1;
        return;
    }
    for (int s = tid; s < ns; s += AT) site_atom[s] = -1;
    __syncthreads();
    const double* S = sites + 3 * (int64_t)s0;
    for (int i = tid; i < n; i += AT) {
        const double px = pos[3 * (int64_t)(n0 + i)], py = pos[3 * (int64_t)(n0 + i) + 1], pz = pos[3 * (int64_t)(n0 + i) + 2];
        double best = 1e300;
        int bs = 0;
        for (int s = 0; s < ns; ++s) {
            const double d = mic_general(Cs, Cis, S[3 * s] - px, S[3 * s + 1] - py, S[3 * s + 2] - pz);
            if (d < best) { best = d; bs = s; }
        }
        atomicMax(&site_atom[bs], i);
    }
    __syncthreads();
    if (tid == 0) {
        int nv = 0;
        for (int s = 0; s < ns; ++s)
            if (site_atom[s] < 0) {
                if (nv < 64) vac[nv] = s;
                ++nv;
            }
        n_vac_s = nv;
    }
    __syncthreads();
    const int nv = n_vac_s;
    if (nv < 1 || nv > max_vac || nv > 64) {
        if (tid == 0) status[g] = 2;
        return;
    }
    const double nn_hi = __dmul_rn(__ddiv_rn(a, sqrt(2.0)), 1.1);
    const int chunk = (ns + AT - 1) / AT;
    for (int v = 0; v < nv; ++v) {
        const double vx = S[3 * vac[v]], vy = S[3 * vac[v] + 1], vz = S[3 * vac[v] + 2];
        const int s_lo = tid * chunk, s_hi = min(ns, s_lo + chunk);
        int mine = 0;
        for (int s = s_lo; s < s_hi; ++s) {
            const double d = mic_general(Cs, Cis, S[3 * s] - vx, S[3 * s + 1] - vy, S[3 * s + 2] - vz);
            if (d > 0.1 && d < nn_hi && site_atom[s] >= 0) ++mine;
        }
        part[tid] = mine;
        __syncthreads();
        if (tid == 0) {
            int run = n_act_s;
            for (int t = 0; t < AT; ++t) { const int c = part[t]; part[t] = run; run += c; }
            part[AT] = run;
        }
        __syncthreads();
        int w = part[tid];
        for (int s = s_lo; s < s_hi; ++s) {
            const double d = mic_general(Cs, Cis, S[3 * s] - vx, S[3 * s + 1] - vy, S[3 * s + 2] - vz);
            if (d > 0.1 && d < nn_hi && site_atom[s] >= 0) {
                const int at = site_atom[s];
                if (w < max_actions) {
                    double* o = actions + ((int64_t)g * max_actions + w) * 4;
                    const double dx = vx - pos[3 * (int64_t)(n0 + at)], dy = vy - pos[3 * (int64_t)(n0 + at) + 1],
                                 dz = vz - pos[3 * (int64_t)(n0 + at) + 2];
                    double f0 = dx * Cis[0] + dy * Cis[3] + dz * Cis[6], f1 = dx * Cis[1] + dy * Cis[4] + dz * Cis[7],
                           f2 = dx * Cis[2] + dy * Cis[5] + dz * Cis[8];
                    f0 -= rint(f0); f1 -= rint(f1); f2 -= rint(f2);
                    o[0] = (double)at;
                    o[1] = f0 * Cs[0] + f1 * Cs[3] + f2 * Cs[6];
                    o[2] = f0 * Cs[1] + f1 * Cs[4] + f2 * Cs[7];
                    o[3] = f0 * Cs[2] + f1 * Cs[5] + f2 * Cs[8];
                }
                ++w;
            }
        }
        __syncthreads();
        if (tid == 0) n_act_s = part[AT];
        __syncthreads();
    }
    if (tid == 0) {
        count[g] = min(n_act_s, max_actions);
        status[g] = n_act_s > max_actions ? 3 : 0;
    }
}

}  // namespace

int launch_action_space_general(rlvd_engine* e, cudaStream_t s, int n_struct, const int* node_ptr, const double* pos, const double* cell,
                                const double* inv_cell, const int* site_ptr, const double* sites, double a, int max_vac, int max_actions,
                                int* count, double* actions, int* status) {
    if (n_struct <= 0) return 0;
    RLVD_CHECK(a > 0.0 && max_vac >= 1 && max_actions >= 1, "rlvd_action_space_general: bad lattice parameter / limits");
    Span sp(e, s, FAM_NEIGHBOR);
    action_space_general_kernel<<<n_struct, AT, 0, s>>>(n_struct, node_ptr, pos, cell, inv_cell, site_ptr, sites, a, max_vac, max_actions,
                                                       count, actions, status);
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

int launch_action_space(rlvd_engine* e, cudaStream_t s, int n_struct, const int* node_ptr, const double* pos, const double* cell,
                        double a, int max_vac, int max_actions, int* count, double* actions, int* status) {
    if (n_struct <= 0) return 0;
    RLVD_CHECK(a > 0.0 && max_vac >= 1 && max_actions >= 1, "rlvd_action_space: bad lattice parameter / limits");
    Span sp(e, s, FAM_NEIGHBOR);
    action_space_kernel<<<n_struct, AT, 0, s>>>(n_struct, node_ptr, pos, cell, a, max_vac, max_actions, count, actions, status);
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rlvd

// ------------------------------------------------------------------------------------------------------------------
// A2/A3 and A5 on the device: the scoring batch of every candidate of every replica is built from the
// replicas' fp64 state without a host round trip, and the chosen hop is applied in place.
// ------------------------------------------------------------------------------------------------------------------
namespace rlvd {

namespace {

// exclusive scan of the candidate counts (single block; a few thousand replicas)
__global__ void __launch_bounds__(1024) candidate_offsets_kernel(int n_rep, const int* __restrict__ count, int* __restrict__ first,
                                                                 int* __restrict__ total) {
    __shared__ int wsum[33];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int carry = 0;
    for (int base = 0; base < n_rep; base += 1024) {
        const int idx = base + tid;
        const int v = idx < n_rep ? count[idx] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int w = wsum[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            wsum[lane] = winc - w;
            if (lane == 31) wsum[32] = winc;
        }
        __syncthreads();
        if (idx < n_rep) first[idx] = carry + wsum[wid] + inc - v;
        carry += wsum[32];
        __syncthreads();
    }
    if (tid == 0) { first[n_rep] = carry; total[0] = carry; }
}

// convert_to_graph_list (rlsim/drl/simulator.py:439-460) + collate (:50-55) for all replicas: candidate c of replica r
// becomes reaction graph g = first[r] + c with reactant structure 2g (a copy of the replica) and product structure
// 2g+1 (pos[atom] += disp, added in fp64 and then rounded to fp32 like ASE positions cast by the collate).
// grid = (total candidates, 2), one CTA per structure.
__global__ void __launch_bounds__(128) expand_candidates_kernel(int n_rep, const int* __restrict__ rep_ptr, const int* __restrict__ first,
                                                                const int* __restrict__ Z64as32, const double* __restrict__ pos,
                                                                const double* __restrict__ cell, const double* __restrict__ actions,
                                                                int max_actions, int n_atoms_per /* uniform replica size */,
                                                                int* __restrict__ node_ptr, int* __restrict__ Zo, float* __restrict__ poso,
                                                                float* __restrict__ cello, float* __restrict__ invo, int* __restrict__ imgo,
                                                                int* __restrict__ rs, int* __restrict__ ps, int* __restrict__ owner,
                                                                const float* __restrict__ inv_in /* [R,3,3] or null */,
                                                                const int* __restrict__ img_in /* [R,3] or null */) {
    const int g = blockIdx.x, side = blockIdx.y, tid = threadIdx.x;
    // replica of graph g: binary search in first[]
    int lo = 0, hi = n_rep;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (first[mid] <= g) lo = mid; else hi = mid;
    }
    const int r = lo, c = g - first[r];
    const int a0 = rep_ptr[r], n = rep_ptr[r + 1] - a0;
    const int st = 2 * g + side;
    const int64_t o0 = (int64_t)st * n_atoms_per;
    const double* act = actions + ((int64_t)r * max_actions + c) * 4;
    const int moved = side == 1 ? (int)act[0] : -1;
    for (int i = tid; i < n; i += 128) {
        double x = pos[3 * (int64_t)(a0 + i)], y = pos[3 * (int64_t)(a0 + i) + 1], z = pos[3 * (int64_t)(a0 + i) + 2];
        if (i == moved) { x = __dadd_rn(x, act[1]); y = __dadd_rn(y, act[2]); z = __dadd_rn(z, act[3]); }
        poso[3 * (o0 + i)] = (float)x;
        poso[3 * (o0 + i) + 1] = (float)y;
        poso[3 * (o0 + i) + 2] = (float)z;
        Zo[o0 + i] = Z64as32[a0 + i];
    }
    if (tid < 9) {
        const double cv = cell[9 * (int64_t)r + tid];
        cello[9 * (int64_t)st + tid] = (float)cv;
        const int row = tid / 3, colm = tid % 3;
        // orthorhombic cells: 1 / L on the diagonal; general cells: the caller's fp32 image of the fp64 inverse
        invo[9 * (int64_t)st + tid] = inv_in != nullptr ? inv_in[9 * (int64_t)r + tid] : (row == colm ? (float)(1.0 / cv) : 0.f);
    }
    if (tid < 3) imgo[3 * (int64_t)st + tid] = img_in != nullptr ? img_in[3 * (int64_t)r + tid] : 0;   // 0: heights >= 2 * cutoff
    if (tid == 0) {
        node_ptr[st] = (int)o0;
        if (side == 0) { rs[g] = st; owner[g] = r; } else ps[g] = st;
        if (g == gridDim.x - 1 && side == 1) node_ptr[st + 1] = (int)(o0 + n_atoms_per);
    }
}

// simulator.py:94-98 for all replicas, then environment.step's position update: p = softmax(Q / kT) in fp32,
// np.random.choice semantics for the draw (cdf = cumsum(p) in fp64, cdf /= cdf[-1], first index with cdf > u), then
// pos[atom] += disp in fp64.  u[r] comes from the caller (np.random.random_sample() for bit-compatible replays, or any
// counter-based generator).
// Counter-based draws: Philox4x32-10 (Salmon et al., SC'11) keyed by the caller's 64-bit seed, counter =
// (replica, step): every (replica, step) pair has its own uniform whatever the launch order, so a trajectory needs no host
// RNG traffic and replays exactly.  u = (x0 * 2^32 + x1) >> 11 scaled to [0, 1) — 53 random bits like numpy's double draw.
__device__ __forceinline__ double philox_uniform(unsigned long long seed, unsigned int replica, unsigned long long step) {
    unsigned int c0 = replica, c1 = (unsigned int)step, c2 = (unsigned int)(step >> 32), c3 = 0u;
    unsigned int k0 = (unsigned int)seed, k1 = (unsigned int)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned int n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const unsigned long long bits = (((unsigned long long)c0 << 32) | (unsigned long long)c1) >> 11;
    return (double)bits * (1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(128) select_apply_kernel(int n_rep, const int* __restrict__ rep_ptr, const int* __restrict__ first,
                                                           const float* __restrict__ rl_q, const float* __restrict__ temperature,
                                                           const double* __restrict__ u, unsigned long long seed, unsigned long long step,
                                                           unsigned int replica0, const double* __restrict__ actions,
                                                           int max_actions, int apply, double* __restrict__ pos,
                                                           int* __restrict__ chosen, int* __restrict__ argmax_out, float* __restrict__ probs,
                                                           float* __restrict__ gamma) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rep) return;
    const int g0 = first[r], nc = first[r + 1] - g0;
    if (nc <= 0) { chosen[r] = -1; argmax_out[r] = -1; if (gamma != nullptr) gamma[r] = 0.f; return; }
    const float kT = temperature[r] * 8.617e-5f;
    if (gamma != nullptr) {                       // TKS total rate: sum(exp(Q / kT)), simulator.py:262
        float gsum = 0.f;
        for (int c = 0; c < nc; ++c) gsum += expf(rl_q[g0 + c] / kT);
        gamma[r] = gsum;
    }
    float m = -3.4e38f;
    int am = 0;
    for (int c = 0; c < nc; ++c) {
        const float z = rl_q[g0 + c] / kT;
        if (z > m) { m = z; am = c; }
    }
    float s = 0.f;
    for (int c = 0; c < nc; ++c) s += expf(rl_q[g0 + c] / kT - m);
    double tot = 0.0;
    for (int c = 0; c < nc; ++c) {
        const float p = expf(rl_q[g0 + c] / kT - m) / s;
        if (probs != nullptr) probs[(int64_t)r * max_actions + c] = p;
        tot += (double)p;
    }
    const double ur = u != nullptr ? u[r] : philox_uniform(seed, replica0 + (unsigned int)r, step);
    double run = 0.0;
    int pick = nc - 1;
    for (int c = 0; c < nc; ++c) {
        run += (double)(expf(rl_q[g0 + c] / kT - m) / s);
        if (run / tot > ur) { pick = c; break; }
    }
    chosen[r] = pick;
    argmax_out[r] = am;
    if (apply) {
        const double* act = actions + ((int64_t)r * max_actions + pick) * 4;
        const int64_t a = rep_ptr[r] + (int)act[0];
        pos[3 * a] += act[1];
        pos[3 * a + 1] += act[2];
        pos[3 * a + 2] += act[3];
    }
}

}  // namespace

// The per-step record of a replica for the rank-0 logger / trajectory writer: one fixed-width fp32 row
// [replica id, n_actions, chosen, E_R, Q[max_actions] (NaN-padded)], packed on the device so that the only exchange of a
// multi-GPU step is ONE all_gather of these rows.
__global__ void pack_records_kernel(int n_rep, const int* __restrict__ first, const float* __restrict__ rl_q, const int* __restrict__ chosen,
                                    const float* __restrict__ E_R, int id0, int id_stride, int max_actions, float* __restrict__ rec) {
    const int r = blockIdx.x, t = threadIdx.x, W = 4 + max_actions;
    if (r >= n_rep) return;
    const int g0 = first[r], nc = first[r + 1] - g0;
    float* row = rec + (int64_t)r * W;
    if (t == 0) {
        row[0] = (float)(id0 + r * id_stride);
        row[1] = (float)nc;
        row[2] = chosen != nullptr ? (float)chosen[r] : -1.f;
        row[3] = E_R != nullptr && nc > 0 ? E_R[g0] : nanf("");
    }
    for (int c = t; c < max_actions; c += blockDim.x) row[4 + c] = c < nc ? rl_q[g0 + c] : nanf("");
}

int launch_pack_records(rlvd_engine* e, cudaStream_t s, int n_rep, const int* first, const float* rl_q, const int* chosen, const float* E_R,
                        int id0, int id_stride, int max_actions, float* rec) {
    if (n_rep <= 0) return 0;
    Span sp(e, s, FAM_READOUT);
    pack_records_kernel<<<n_rep, 32, 0, s>>>(n_rep, first, rl_q, chosen, E_R, id0, id_stride, max_actions, rec);
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

int launch_candidate_offsets(rlvd_engine* e, cudaStream_t s, int n_rep, const int* count, int* first, int* total) {
    Span sp(e, s, FAM_NEIGHBOR);
    candidate_offsets_kernel<<<1, 1024, 0, s>>>(n_rep, count, first, total);
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

int launch_expand_candidates(rlvd_engine* e, cudaStream_t s, int n_rep, int n_cand, const int* rep_ptr, const int* first, const int* Z,
                             const double* pos, const double* cell, const double* actions, int max_actions, int n_atoms_per,
                             int* node_ptr, int* Zo, float* poso, float* cello, float* invo, int* imgo, int* rs, int* ps, int* owner,
                             const float* inv_in, const int* img_in) {
    if (n_cand <= 0) return 0;
    Span sp(e, s, FAM_NEIGHBOR);
    expand_candidates_kernel<<<dim3(n_cand, 2), 128, 0, s>>>(n_rep, rep_ptr, first, Z, pos, cell, actions, max_actions, n_atoms_per,
                                                            node_ptr, Zo, poso, cello, invo, imgo, rs, ps, owner, inv_in, img_in);
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

int launch_select_apply(rlvd_engine* e, cudaStream_t s, int n_rep, const int* rep_ptr, const int* first, const float* rl_q,
                        const float* temperature, const double* u, unsigned long long seed, unsigned long long step, unsigned int replica0,
                        const double* actions, int max_actions, int apply, double* pos,
                        int* chosen, int* argmax_out, float* probs, float* gamma) {
    if (n_rep <= 0) return 0;
    Span sp(e, s, FAM_READOUT);
    select_apply_kernel<<<(n_rep + 127) / 128, 128, 0, s>>>(n_rep, rep_ptr, first, rl_q, temperature, u, seed, step, replica0, actions, max_actions, apply,
                                                           pos, chosen, argmax_out, probs, gamma);
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rlvd

// ------------------------------------------------------------------------------------------------------------------
// Warren–Cowley pair counts (rlsim/utils/sro.py:49-83, called per candidate at simulator.py:60-92):
// NN_dist = smallest minimum-image distance above 0.1, r_cut = (1 + sqrt2)/2 * NN_dist, ordered pair counts per
// species pair inside (0.1, r_cut).  One CTA per structure (orthorhombic cells), fp64 distances in numpy's operation
// order; integer counts, so the host formula gives the reference's matrix exactly.
// ------------------------------------------------------------------------------------------------------------------
namespace rlvd {

namespace {

constexpr int SRO_MAX_SPECIES = 8;

__global__ void __launch_bounds__(256) sro_counts_kernel(int n_struct, const int* __restrict__ node_ptr, const double* __restrict__ pos,
                                                         const double* __restrict__ cell, const int* __restrict__ sid, int n_species,
                                                         long long* __restrict__ counts, int* __restrict__ status) {
    __shared__ double red[256];
    __shared__ unsigned long long cnt[SRO_MAX_SPECIES * SRO_MAX_SPECIES];
    const int g = blockIdx.x, tid = threadIdx.x;
    if (g >= n_struct) return;
    const int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
    const double* C = cell + 9 * (int64_t)g;
    const bool diag = C[1] == 0.0 && C[2] == 0.0 && C[3] == 0.0 && C[5] == 0.0 && C[6] == 0.0 && C[7] == 0.0;
    if (!diag) {
        if (tid == 0) status[g] = 1;
        return;
    }
    const double Lx = C[0], Ly = C[4], Lz = C[8];
    if (tid < SRO_MAX_SPECIES * SRO_MAX_SPECIES) cnt[tid] = 0ull;
    // ---- pass 1: nearest-neighbour distance
    double best = 1e300;
    for (int64_t p = tid; p < (int64_t)n * n; p += 256) {
        const int i = (int)(p / n), j = (int)(p % n);
        const double* pi = pos + 3 * (int64_t)(n0 + i);
        const double* pj = pos + 3 * (int64_t)(n0 + j);
        const double d = norm3(mic1(__dsub_rn(pj[0], pi[0]), Lx), mic1(__dsub_rn(pj[1], pi[1]), Ly), mic1(__dsub_rn(pj[2], pi[2]), Lz));
        if (d > 0.1 && d < best) best = d;
    }
    red[tid] = best;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) red[tid] = fmin(red[tid], red[tid + o]);
        __syncthreads();
    }
    const double r_cut = __dmul_rn(__ddiv_rn(__dadd_rn(1.0, sqrt(2.0)), 2.0), red[0]);
    // ---- pass 2: ordered pair counts per species pair
    unsigned int mine[SRO_MAX_SPECIES * SRO_MAX_SPECIES];
#pragma unroll
    for (int t = 0; t < SRO_MAX_SPECIES * SRO_MAX_SPECIES; ++t) mine[t] = 0u;
    for (int64_t p = tid; p < (int64_t)n * n; p += 256) {
        const int i = (int)(p / n), j = (int)(p % n);
        const double* pi = pos + 3 * (int64_t)(n0 + i);
        const double* pj = pos + 3 * (int64_t)(n0 + j);
        const double d = norm3(mic1(__dsub_rn(pj[0], pi[0]), Lx), mic1(__dsub_rn(pj[1], pi[1]), Ly), mic1(__dsub_rn(pj[2], pi[2]), Lz));
        if (d > 0.1 && d < r_cut) {
            const int k = sid[n0 + i] * n_species + sid[n0 + j];
#pragma unroll
            for (int t = 0; t < SRO_MAX_SPECIES * SRO_MAX_SPECIES; ++t) mine[t] += (t == k) ? 1u : 0u;
        }
    }
    for (int t = 0; t < n_species * n_species; ++t)
        if (mine[t]) atomicAdd(&cnt[t], (unsigned long long)mine[t]);      // integer sums: order-independent
    __syncthreads();
    if (tid < n_species * n_species) counts[(int64_t)g * n_species * n_species + tid] = (long long)cnt[tid];
    if (tid == 0) status[g] = 0;
}

}  // namespace

int launch_sro_counts(rlvd_engine* e, cudaStream_t s, int n_struct, const int* node_ptr, const double* pos, const double* cell,
                      const int* sid, int n_species, long long* counts, int* status) {
    if (n_struct <= 0) return 0;
    RLVD_CHECK(n_species >= 1 && n_species <= SRO_MAX_SPECIES, "rlvd_sro_counts: n_species %d not in [1,%d]", n_species, SRO_MAX_SPECIES);
    Span sp(e, s, FAM_NEIGHBOR);
    sro_counts_kernel<<<n_struct, 256, 0, s>>>(n_struct, node_ptr, pos, cell, sid, n_species, counts, status);
    sp.end(1);
    RLVD_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rlvd
