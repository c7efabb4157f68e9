// Symbolic phase (host, once per multigrid level): everything structural that the reference redoes
// on every Newton step (spdiagm + structural hash in amgb_diag, reference
// src/MultiGridBarrierMPI.jl:137-147 and tools/profile_hash.jl:41-66; SpGEMM symbolic of
// D_j' * diag * D_k and R' * H * R, test/test_map_rows_compare.jl:102-123,165-171) is computed here
// exactly once and frozen into index arrays the numeric kernels replay.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mgb {

struct HostCSR {
    int64_t nrows = 0, ncols = 0;
    std::vector<int64_t> ptr;
    std::vector<int32_t> idx;
    std::vector<double> val;
    int64_t nnz() const { return (int64_t)idx.size(); }
};

// C = A * B, sorted columns, keeps structural zeros (Julia spmatmul semantics).
HostCSR spgemm(const HostCSR& A, const HostCSR& B);
HostCSR transpose(const HostCSR& A);
// nnz of the structural pattern of sum_{j,k} D_j' * diag * D_k
int64_t count_gram_pattern(const std::vector<HostCSR>& D);

// Layout of the per-element slot record written by the element kernel and replayed by the gather
// kernel.  Shared by host symbolic code and device code (see kernels.cuh: must stay in sync).
struct SlotLayout {
    int B = 0, LPE = 0, NU = 0, dim = 0;
    bool slack = false, fine = false;
    bool mma = false;   // coarse fem2d levels on the tensor cores: full 8 x 8 blocks uu | us | ss (ElemParams::mma)
    int off_uu = 0, off_us = 0, off_ss = 0, off_ut = 0, off_st = 0, off_tt = 0;
    int NS = 0;  // doubles per element
    void build(int B_, int dim_, bool slack_, bool fine_, bool mma_ = false);
    int tri(int q, int q2) const;  // packed upper-triangle index, q <= q2 < B
    // slot of packed entry pk of a block of padded size npad that went through the transposing
    // butterfly: lane l ends with entries [l*K,(l+1)*K) and stores entry r at  off + r*LPE + l
    int packed(int off, int pk, int npad) const { const int K = npad / LPE; return off + (pk % K) * LPE + pk / K; }
    int ntri_pad() const;
    int nfull_pad() const;
};

struct ElementPlan {
    bool ok = false;       // element-block structure detected and supported by the fused kernels
    std::string why;       // reason when !ok
    int B = 0, LPE = 0, dim = 0, NU = 0, ND = 0;
    bool slack = false, fine = false;   // slack: three-variable table [u.id; u.d*; v1.id; v2.id] (modes 1 and 2)
    int mode = 0;                       // 0 one cone, 1 feasibility (cone on s + tau, -log(1+tau)), 2 two cones (parabolic)
    int64_t E = 0, nloc = 0, m = 0;
    // ---- dense path (kernels_dense.cuh): elements with more than 8 nodes whose level operators have long rows - the
    // coarse levels of fem3d's Q3 hexahedra, where every fine point touches up to 64 unknowns per variable and the
    // product lists of the CSR path explode.  Consecutive fine elements are grouped (greedily, <= DENSE_NB dofs per
    // variable: the children of one coarse element), the operators are stored as DENSE rows over the group's dofs, and
    // a group's points are cut into chunks; one CTA contracts a chunk into full uu / us / ss blocks
    // (sel record = 3 * DENSE_NB^2 doubles per chunk), which the ordinary gather replays into R'HR.
    bool dense = false;
    int64_t d_ngroups = 0, d_nchunks = 0;
    std::vector<int32_t> d_gdof;   // [group][2][DENSE_NB] global dof or -1
    std::vector<int64_t> d_chunk;  // [chunk][3] = {group, first point, end point} (local row ids)
    std::vector<double> d_rows;    // [point][dim + 2][DENSE_STRIDE]: derivative rows (u), u.id row, s.id row over the group's dofs
    int agg = 1;                   // coarse levels: aligned groups of `agg` consecutive elements share all their dofs (children of
                                   // one coarse element): one slot / gradient record per group (kernels.cuh agg_reduce)
    bool mma = false;              // coarse fem2d level contracted on the tensor cores (SlotLayout::mma)
    int64_t out0 = 0, m_out = 0;   // output rows [out0, out0 + m_out) of the m unknowns (whole range unless sharded)
    SlotLayout lay;
    std::vector<int32_t> lcols;    // [E][NU][LPE]  global dof or -1
    int RW = 0;                    // doubles per point record (even)
    std::vector<double> prec;      // [nloc][RW]: derivative rows (dim*B), w, then fine: own_val[NU] + packed
                                   // own_lq bytes (255 = none); coarse: dense id-like rows [NU][B]
    // fixed output pattern + replay lists
    std::vector<int32_t> h_rowptr, h_colidx;  // m_out+1 (row a - out0), nnzH (global column ids)
    std::vector<int64_t> h_cptr;              // nnzH+1
    std::vector<int32_t> h_cidx;              // contribution -> e*NS + slot
    std::vector<int64_t> g_cptr;              // m_out+1
    std::vector<int32_t> g_cidx;              // contribution -> (e*NU+v)*LPE + q
};

struct BarrierDesc {
    int kind = 1, nidx = 0, idx[8] = {0};
    double p = 1.0;
    int slack = 0;
    int nidx2 = 0, idx2[8] = {0};  // optional second cone (intersection)
    double p2 = 2.0;
};

// D: nD operators restricted to the local rows (nloc x N), R: N x m.
// Tries to detect the broken-element block structure the fused kernels exploit.
// out0/out1: keep only the output rows (unknowns) in [out0, out1) (out1 < 0: all m) - a sharded plan whose local
// quadrature rows contain every element touching those unknowns completes them without any exchange.
void build_element_plan(const std::vector<HostCSR>& D, const HostCSR& R, int64_t n_global, const double* w_local,
                        const BarrierDesc& bar, ElementPlan& out, bool want_hessian = true, int64_t out0 = 0, int64_t out1 = -1,
                        bool allow_agg = true);


// ---- multi-GPU (one process per GPU): owner-computes sharding ---------------------------------------
// The outputs (gradient entries, rows of R'HR) are split in contiguous blocks of the m unknowns
// (HPCSparseArrays row partition).  A rank evaluates every broken element that touches one of
// its rows, so every owned row is completed locally: no Hessian or gradient value crosses NVLink.  Elements
// on the interface are evaluated by up to two (P = 2) ... a few ranks; that redundancy costs microseconds of
// arithmetic and replaces the exchange of 50-94 % of all element contributions that the [u dofs | s dofs]
// numbering of R = blockdiag(R_u, R_s) would force.  Only the three objective scalars
// are summed across ranks (peer-memory words with embedded epoch flags, kernels_dist.cuh).
constexpr int DIST_MAX_RANKS = 16;
constexpr int DENSE_NB = 64;       // dofs per variable of a dense group (Q3 hexahedron: 4^3 nodes)
constexpr int DENSE_STRIDE = 68;   // doubles between consecutive dense rows (64 + 4 padding: bank-conflict-free MMA fragments)
constexpr int DENSE_CHUNK = 512;   // points per chunk (a multiple of the kernel's point tile)

// Quadrature rows (whole elements) rank `rank` evaluates: every element with a dof in its output block
// [out_part[rank], out_part[rank+1]).  The objective scalars of an element are counted by exactly one of the ranks
// that evaluate it anyway - the owner of its first dof of the last state variable (s is never eliminated; any dof /
// rank 0 as fall-backs) - and those "primary" elements are listed first.  lcols: element -> dof table of the global
// element plan ([E][NU][LPE]).
void dist_select_elements(const std::vector<int32_t>& lcols, int64_t E, int NU, int LPE, int B, int rank, int nranks,
                          const int64_t* out_part, std::vector<int64_t>& rows_sel, int64_t& n_primary_rows);

struct CsrPlan {
    int ND = 0, NU = 0;
    int64_t nloc = 0, m = 0;
    std::vector<HostCSR> E;   // E_k = D_k R   (nloc x m)
    // gradient: all E_k' merged into one list per unknown a:  g[a] = sum_r gt_coef[r] * gy[gt_src[r]],
    // gt_src = k*nloc + i, listed in the fixed order (k, i)
    std::vector<int32_t> gt_ptr, gt_src;
    std::vector<double> gt_coef;
    std::vector<int32_t> h_rowptr, h_colidx;
    // V = w .* F2 is symmetric and structurally zero outside the cones' index sets: only the unique pairs
    // (ka <= kb) that some cone couples get a column; pair_col[ka][kb] = column or -1
    int npair = 0;
    int pair_a[36] = {0}, pair_b[36] = {0};
    int pair_col[8][8];
    // numeric-only replay of sum_jk E_j' diag(V_jk) E_k on the frozen (symmetric) pattern, upper triangle only:
    // upper entry j (row a <= column b) sums  coef[r] * V[vsrc[r]],  r in [prod_ptr[j], prod_ptr[j+1]),
    // coef = E_ka[i,a]*E_kb[i,b], vsrc = pair_col[ka][kb]*nloc + i, in the fixed order (ka, i, kb); the value
    // goes to position up_t[j] and to its mirror up_m[j] (-1 on the diagonal)
    std::vector<int32_t> up_t, up_m, prod_ptr;
    std::vector<double> prod_coef;
    std::vector<int32_t> prod_v;
    int32_t max_row = 0;
};

void build_csr_plan(const std::vector<HostCSR>& D, const HostCSR& R, const BarrierDesc& bar, CsrPlan& out,
                    bool want_hessian = true);

}  // namespace mgb
