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

}  // namespace

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
