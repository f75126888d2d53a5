// Einsum planner (host) — canonicalises a subscript string into one of the device kernels.
// Replaces `np.einsum(self.op_str, *arr)` at reference autodiff/core/ops.py:180; the subscript grammar
// follows ops.py:131,167-170 (explicit "->", single-letter indices; "..." is expanded by the host
// binding before it gets here).
//
//   every letter gets an extent and one stride per operand (0 when the operand does not carry it, the
//   SUM of strides when it carries it twice: diagonals)  =>  out[free] = sum_summed prod_t op_t[...]
//
//   no summed letter                      -> ewise product program (broadcast / outer / Hadamard / permute)
//   2 operands, letters split cleanly into batch | M | N | K groups that each collapse to one stride
//   and M,N >= 16, K >= 4                -> strided-batched FP64 GEMM (TMA + DMMA)
//   3+ operands with summed letters, large  -> chain of pairwise contractions (each again one of these forms)
//   anything else                         -> sum-of-products reduction kernel (row or column mode)
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include <stdlib.h>

#include "adb_internal.h"

namespace adb {

struct Letter {
    char c;
    long long extent;
    bool in_out;
    long long os;                       // output stride (0 when summed)
    long long s[ADB_MAX_OPERANDS];      // per-operand stride (0 = absent / broadcast)
    bool present[ADB_MAX_OPERANDS];     // operand really varies along this letter
};

struct Problem {
    int n_ops;
    std::vector<Letter> letters;        // extent > 1 only
    const double* op[ADB_MAX_OPERANDS];
    double* out;
    bool empty_out;                     // some free extent is 0
    bool empty_sum;                     // some summed extent is 0
    adb_tensor out_t;                   // original output view (for fills)
};

static int parse(const char* op_str, int n_ops, const adb_tensor* ops, const adb_tensor* out, Problem& P) {
    if (!op_str) return set_error("einsum: null subscripts");
    if (n_ops < 1 || n_ops > ADB_MAX_OPERANDS) return set_error("einsum: %d operands (max %d)", n_ops, ADB_MAX_OPERANDS);
    const char* arrow = strstr(op_str, "->");
    if (!arrow) return set_error("einsum: subscripts '%s' need an explicit '->' (reference ops.py:131)", op_str);
    std::vector<std::string> subs;
    {
        std::string cur;
        for (const char* p = op_str; p < arrow; ++p) {
            if (*p == ',') { subs.push_back(cur); cur.clear(); }
            else if (*p == ' ') continue;
            else cur.push_back(*p);
        }
        subs.push_back(cur);
    }
    std::string osub;
    for (const char* p = arrow + 2; *p; ++p)
        if (*p != ' ') osub.push_back(*p);
    if ((int)subs.size() != n_ops) return set_error("Number of operands doesn't match the einsum string!");
    P.n_ops = n_ops;
    P.letters.clear();
    P.empty_out = P.empty_sum = false;
    P.out = out->ptr;
    P.out_t = *out;
    std::vector<Letter> all;
    auto find = [&](char c) -> Letter* {
        for (auto& l : all) if (l.c == c) return &l;
        return nullptr;
    };
    for (int t = 0; t < n_ops; ++t) {
        P.op[t] = ops[t].ptr;
        if ((int)subs[t].size() != ops[t].rank)
            return set_error("einsum: operand %d has rank %d but subscripts '%s'", t, ops[t].rank, subs[t].c_str());
        for (int k = 0; k < ops[t].rank; ++k) {
            const char c = subs[t][k];
            if (!((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'))) return set_error("einsum: bad subscript character '%c'", c);
            Letter* l = find(c);
            if (!l) {
                Letter nl;
                memset(&nl, 0, sizeof nl);
                nl.c = c;
                nl.extent = 1;
                all.push_back(nl);
                l = &all.back();
            }
            const long long d = ops[t].shape[k];
            if (d != 1) {
                if (l->extent != 1 && l->extent != d)
                    return set_error("einsum: operand %d dim %d has extent %lld but '%c' is already %lld", t, k, d, c, l->extent);
                l->extent = d;
                l->s[t] += ops[t].stride[k];
                l->present[t] = true;
            }
        }
    }
    if ((int)osub.size() != out->rank) return set_error("einsum: output has rank %d but subscripts '%s'", out->rank, osub.c_str());
    for (int k = 0; k < out->rank; ++k) {
        const char c = osub[k];
        Letter* l = find(c);
        if (!l) return set_error("einsum: output subscript '%c' never appeared in an input", c);
        if (l->in_out) return set_error("einsum: output subscript '%c' appears twice", c);
        l->in_out = true;
        l->os = out->stride[k];
        if (out->shape[k] != l->extent)
            return set_error("einsum: output dim %d has extent %lld, expected %lld", k, (long long)out->shape[k], l->extent);
    }
    for (auto& l : all) {
        if (l.extent == 0) { (l.in_out ? P.empty_out : P.empty_sum) = true; }
        if (l.extent > 1) P.letters.push_back(l);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
static int run_fill_zero(const Problem& P, cudaStream_t st, std::string* desc) {
    if (desc) { *desc = "fill0"; return 0; }
    adb_ew_instr code[2] = {{ADB_EW_LOAD | ADB_EW_SRC_IMM, 0, 0.0}, {ADB_EW_STORE_OUT, 0, 0.0}};
    return ewise_run(code, 2, 0, nullptr, 1, &P.out_t, st);
}

static int run_product(const Problem& P, cudaStream_t st, std::string* desc) {
    // iteration space = free letters ordered by output stride (outermost first)
    std::vector<const Letter*> ls;
    for (auto& l : P.letters) ls.push_back(&l);
    std::sort(ls.begin(), ls.end(), [](const Letter* a, const Letter* b) { return a->os > b->os; });
    const int rank = (int)ls.size();
    if (rank > ADB_MAX_RANK) return set_error("einsum: %d free letters (max %d)", rank, ADB_MAX_RANK);
    if (desc) {
        char b[128];
        snprintf(b, sizeof b, "ewise product ops=%d rank=%d", P.n_ops, rank);
        *desc = b;
        return 0;
    }
    adb_tensor o;
    adb_tensor in[ADB_MAX_OPERANDS];
    o.ptr = P.out; o.rank = rank;
    for (int k = 0; k < rank; ++k) { o.shape[k] = ls[k]->extent; o.stride[k] = ls[k]->os; }
    adb_ew_instr code[ADB_MAX_OPERANDS + 1];
    for (int t = 0; t < P.n_ops; ++t) {
        in[t].ptr = const_cast<double*>(P.op[t]);
        in[t].rank = rank;
        for (int k = 0; k < rank; ++k) {
            in[t].shape[k] = ls[k]->present[t] ? ls[k]->extent : 1;
            in[t].stride[k] = ls[k]->present[t] ? ls[k]->s[t] : 0;
        }
        code[t].op = (t == 0 ? ADB_EW_LOAD : ADB_EW_MUL) | ADB_EW_SRC_IN;
        code[t].arg = t;
        code[t].imm = 0.0;
    }
    code[P.n_ops].op = ADB_EW_STORE_OUT; code[P.n_ops].arg = 0; code[P.n_ops].imm = 0.0;
    return ewise_run(code, P.n_ops + 1, P.n_ops, in, 1, &o, st);
}

// Collapse a group of letters to (extent, stride per tensor). `key` orders them outermost-first.
struct Group {
    long long extent = 1;
    long long sa = 0, sb = 0, sc = 0;
    bool ok = true;
    int count = 0;
};

static Group collapse_group(std::vector<const Letter*> ls, bool useA, bool useB, bool useC, int ia, int ib) {
    Group g;
    g.count = (int)ls.size();
    if (ls.empty()) return g;
    auto key = [&](const Letter* l) { return useC ? l->os : l->s[ia]; };
    std::sort(ls.begin(), ls.end(), [&](const Letter* a, const Letter* b) { return key(a) > key(b); });
    const Letter* in = ls.back();
    g.extent = in->extent;
    g.sa = useA ? in->s[ia] : 0;
    g.sb = useB ? in->s[ib] : 0;
    g.sc = useC ? in->os : 0;
    for (int i = (int)ls.size() - 2; i >= 0; --i) {
        const Letter* l = ls[i];
        if (useA && l->s[ia] != g.sa * g.extent) g.ok = false;
        if (useB && l->s[ib] != g.sb * g.extent) g.ok = false;
        if (useC && l->os != g.sc * g.extent) g.ok = false;
        g.extent *= l->extent;
    }
    return g;
}

static bool try_gemm(const Problem& P, adb_gemm_desc& d, bool* swapped = nullptr) {
    if (P.n_ops != 2) return false;
    std::vector<const Letter*> B_, M_, N_, K_;
    for (auto& l : P.letters) {
        const bool a = l.present[0], b = l.present[1], o = l.in_out;
        if (a && b && o) B_.push_back(&l);
        else if (a && !b && o) M_.push_back(&l);
        else if (!a && b && o) N_.push_back(&l);
        else if (a && b && !o) K_.push_back(&l);
        else return false;  // a letter summed out of a single operand: not a plain GEMM
    }
    if (M_.empty() || N_.empty() || K_.empty()) return false;
    Group gm = collapse_group(M_, true, false, true, 0, 1);
    Group gn = collapse_group(N_, false, true, true, 0, 1);
    Group gk = collapse_group(K_, true, true, false, 0, 1);
    Group gb = collapse_group(B_, true, true, true, 0, 1);
    if (!gm.ok || !gn.ok || !gk.ok || !gb.ok) return false;
    // GEMM needs a tile's worth of rows and columns - except tall-skinny problems (dense layers onto a few classes,
    // im2col convolutions onto a few channels), which the 256x32 tile configuration of gemm_f64.cu handles well
    const long long mn_min = gm.extent < gn.extent ? gm.extent : gn.extent, mn_max = gm.extent < gn.extent ? gn.extent : gm.extent;
    if (gk.extent < 4 || mn_min < 4 || (mn_min < 16 && mn_max < 256)) return false;
    if (gm.sa < 0 || gk.sa < 0 || gk.sb < 0 || gn.sb < 0 || gm.sc < 0 || gn.sc < 0) return false;
    d.batch = gb.extent;
    d.K = gk.extent;
    if (swapped) *swapped = !(gn.sc <= gm.sc);
    if (gn.sc <= gm.sc) {
        d.M = gm.extent; d.N = gn.extent;
        d.A = P.op[0]; d.a_sm = gm.sa; d.a_sk = gk.sa; d.a_sb = gb.sa;
        d.B = P.op[1]; d.b_sk = gk.sb; d.b_sn = gn.sb; d.b_sb = gb.sb;
        d.C = P.out; d.c_sm = gm.sc; d.c_sn = gn.sc; d.c_sb = gb.sc;
    } else {  // output is "N-major": compute the transpose so the fast C stride follows the GEMM n index
        d.M = gn.extent; d.N = gm.extent;
        d.A = P.op[1]; d.a_sm = gn.sb; d.a_sk = gk.sb; d.a_sb = gb.sb;
        d.B = P.op[0]; d.b_sk = gk.sa; d.b_sn = gm.sa; d.b_sb = gb.sa;
        d.C = P.out; d.c_sm = gn.sc; d.c_sn = gm.sc; d.c_sb = gb.sc;
    }
    return true;
}

static int run_reduce(const Problem& P, cudaStream_t st, std::string* desc) {
    if (P.n_ops > ADB_RED_MAX_OPS) return set_error("einsum: %d operands with summed letters (max %d)", P.n_ops, ADB_RED_MAX_OPS);
    // the operand spanning the most elements decides the memory-friendly ordering
    int big = 0;
    double big_elems = -1;
    for (int t = 0; t < P.n_ops; ++t) {
        double e = 1;
        for (auto& l : P.letters) if (l.present[t]) e *= (double)l.extent;
        if (e > big_elems) { big_elems = e; big = t; }
    }
    auto stride_key = [&](const Letter* l) -> long long {
        if (l->present[big]) return l->s[big] < 0 ? -l->s[big] : l->s[big];
        return (1LL << 60) + (l->os < 0 ? -l->os : l->os);   // letters the big operand lacks go outermost
    };
    std::vector<const Letter*> fr, rd;
    for (auto& l : P.letters) (l.in_out ? fr : rd).push_back(&l);
    std::sort(fr.begin(), fr.end(), [&](const Letter* a, const Letter* b) { return stride_key(a) < stride_key(b); });
    std::sort(rd.begin(), rd.end(), [&](const Letter* a, const Letter* b) { return stride_key(a) < stride_key(b); });

    ReduceSpec S;
    memset(&S, 0, sizeof S);
    S.n_ops = P.n_ops;
    for (int t = 0; t < P.n_ops; ++t) S.op[t] = P.op[t];
    S.out = P.out;
    // collapse innermost-first lists
    auto push = [&](const std::vector<const Letter*>& ls, bool is_out, long long* dim, long long (*ops)[4], long long* outs, int& n) -> bool {
        n = 0;
        for (const Letter* l : ls) {
            bool merged = false;
            if (n > 0) {
                bool ok = true;
                for (int t = 0; t < P.n_ops && ok; ++t) ok = (l->present[t] ? l->s[t] : 0) == ops[t][n - 1] * dim[n - 1];
                if (is_out && ok) ok = l->os == outs[n - 1] * dim[n - 1];
                if (ok) { dim[n - 1] *= l->extent; merged = true; }
            }
            if (!merged) {
                if (n == 4) return false;
                dim[n] = l->extent;
                for (int t = 0; t < P.n_ops; ++t) ops[t][n] = l->present[t] ? l->s[t] : 0;
                if (is_out) outs[n] = l->os;
                ++n;
            }
        }
        return true;
    };
    if (!push(fr, true, S.odim, S.op_os, S.out_s, S.n_o)) return set_error("einsum: more than 4 non-collapsible output dims");
    if (!push(rd, false, S.rdim, S.op_rs, nullptr, S.n_r)) return set_error("einsum: more than 4 non-collapsible summed dims");

    // row mode when the big operand's fastest letter is summed, column mode when it is free
    long long min_free = 1LL << 62, min_red = 1LL << 62;
    for (const Letter* l : fr) if (l->present[big]) min_free = std::min(min_free, stride_key(l));
    for (const Letter* l : rd) if (l->present[big]) min_red = std::min(min_red, stride_key(l));
    S.group = 1;
    if (min_red < min_free && S.n_r > 0) {
        if (S.rdim[0] >= 64) S.group = 32;
        else if (S.rdim[0] >= 8) S.group = 8;
    }
    if (desc) {
        char b[160];
        long long O = 1, R = 1;
        for (int k = 0; k < S.n_o; ++k) O *= S.odim[k];
        for (int k = 0; k < S.n_r; ++k) R *= S.rdim[k];
        snprintf(b, sizeof b, "reduce ops=%d G=%d O=%lld R=%lld odims=%d rdims=%d", P.n_ops, S.group, O, R, S.n_o, S.n_r);
        *desc = b;
        return 0;
    }
    return reduce_run(S, st);
}

// ---- ragged edge of a GEMM ------------------------------------------------------------------------------------------
// The NN bias trick (reference high_level_ops.py:32-35) makes one GEMM dimension 4097: a 33rd (65th) column of tiles that
// is 1/128 (1/64) full.  With ~7-14 tiles per SM that extra column usually costs a whole extra WAVE: at 4097 x 4096 x 8192
// the busiest SMs run 8 tiles of 128x128 instead of 7 (+14 %; ncu on the 1024-row shard: SMs idle 14 % of the launch).
// When cutting the <= 4 ragged rows/columns off saves more than they cost as a bandwidth-bound reduction (one pass over
// the other operand), the contraction runs as  GEMM on the aligned part  +  run_reduce on the edge.
// Time model: 64x64 k-tile units per SM (a 128x128 tile = 4 units), tile choice as in gemm_f64.cu.
static double gemm_makespan_units(long long M, long long N) {
    const long long big = ((M + 127) / 128) * ((N + 127) / 128);
    const double waste_big = (double)((big + 147) / 148 * 148) / (double)big;
    if (waste_big <= 1.05) return 4.0 * (double)((big + 147) / 148);
    const long long small = ((M + 63) / 64) * ((N + 63) / 64);
    return (double)((small + 147) / 148);
}

// Which letter (index into P.letters) to cut and by how much; -1 = no split.
static int ragged_edge(const Problem& P, const adb_gemm_desc& d, bool swapped, long long* edge_len) {
    if (d.batch != 1 || d.K < 512 || g_precision != 0) return -1;
    if (const char* e = getenv("ADB_EDGE_SPLIT")) if (e[0] == '0') return -1;      // A/B hook (tools/bench_edge_split.py)
    if (d.N <= 32 || d.M <= 32) return -1;                    // tall-skinny problems have their own tile configurations
    const double unit_s = 64.0 * 64.0 * (double)d.K * 2.0 / (37.0e12 / 148.0);      // one 64x64xK tile on one SM
    const double full = gemm_makespan_units(d.M, d.N);
    int best = -1;
    double best_gain = 0.0;
    for (int side = 0; side < 2; ++side) {                    // 0: letters only operand 0 and the output carry, 1: operand 1
        int idx = -1, count = 0;
        for (int i = 0; i < (int)P.letters.size(); ++i) {
            const Letter& l = P.letters[i];
            if (l.in_out && l.present[side] && !l.present[1 - side]) { idx = i; ++count; }
        }
        if (count != 1) continue;
        const bool is_m = (side == 0) != swapped;             // does this letter play the GEMM's M role?
        // Measured (tools/bench_edge_split.py, profiles/r01_edge_split_ab.log): ragged COLUMNS: dX of a 4096-wide layer
        // 1.117 -> 0.997 ms at 1024 rows, 7.85 -> 7.71 ms at 8192; ragged ROWS (dW, M = 4097): 8.19 -> 7.73 ms at 8192 rows,
        // 1.027 -> 1.009 ms at 1024 (with the 8-box tensor maps of the first version the aligned MN-contiguous GEMM itself was
        // the slow part and the cut lost: 8.29 -> 8.52 ms).  ADB_EDGE_SPLIT=3 restricts the cut to columns again.
        static const bool cols_only = getenv("ADB_EDGE_SPLIT") && getenv("ADB_EDGE_SPLIT")[0] == '3';
        if (is_m && cols_only) continue;
        const long long X = P.letters[idx].extent, other = is_m ? d.N : d.M;
        if (X != (is_m ? d.M : d.N)) continue;
        const long long r = X % 128;
        if (r == 0 || r > 4 || X < 1024) continue;
        const double main_units = is_m ? gemm_makespan_units(X - r, d.N) : gemm_makespan_units(d.M, X - r);
        const double edge_s = (double)r * 8.0 * (double)other * (double)d.K / 5.0e12 + 10e-6;
        const double gain = (full - main_units) * unit_s - 1.5 * edge_s;
        if (gain > best_gain) { best_gain = gain; best = idx; *edge_len = r; }
    }
    return best;
}

static int run_problem(const Problem& P, cudaStream_t st, std::string* desc);

// ---- three or more operands with summed letters: a chain of pairwise contractions ---------------------------------------
// np.einsum (reference ops.py:180, optimize=False) walks the full iteration space  prod(extent of every letter)  with one
// multiply per operand; so did the sum-of-products reduction kernel here (on the DFMA pipe).  Contracting two operands at
// a time - greedily the pair with the smallest iteration space, summing every letter nothing else needs - turns
// 'ij,jk,kl->il' from I*J*K*L products into two GEMMs on the DMMA pipe, and the gradient of a 3-operand einsum (again a
// 3-operand einsum, ops.py:182-213) likewise.  Intermediates are dense stream-ordered temporaries laid out
// [letters of both | letters of the first | letters of the second] so the pair step itself collapses to a GEMM.
struct PairPlan {
    int i, j;
    double iter, inter;          // iteration space of the pair step, elements of its intermediate
};

static bool pair_letter_kept(const Problem& P, const Letter& l, int i, int j) {
    if (l.in_out) return true;
    for (int t = 0; t < P.n_ops; ++t)
        if (t != i && t != j && l.present[t]) return true;
    return false;
}

static PairPlan best_pair(const Problem& P) {
    PairPlan best{-1, -1, 0, 0};
    for (int i = 0; i < P.n_ops; ++i)
        for (int j = i + 1; j < P.n_ops; ++j) {
            double iter = 1, inter = 1;
            for (auto& l : P.letters) {
                if (!l.present[i] && !l.present[j]) continue;
                iter *= (double)l.extent;
                if (pair_letter_kept(P, l, i, j)) inter *= (double)l.extent;
            }
            if (best.i < 0 || iter < best.iter || (iter == best.iter && inter < best.inter)) best = PairPlan{i, j, iter, inter};
        }
    return best;
}

// Is the chain worth it?  Simulates the greedy order on extents only.
static bool pairwise_pays(const Problem& P0) {
    static int pref = -1;                // ADB_EINSUM_PAIRWISE=0: off (A/B measurements)
    if (pref < 0) { const char* e = getenv("ADB_EINSUM_PAIRWISE"); pref = (e && e[0] == '0') ? 0 : 1; }
    if (!pref) return false;
    double naive = 1, biggest = 1;
    for (auto& l : P0.letters) naive *= (double)l.extent;
    for (int t = 0; t <= P0.n_ops; ++t) {                       // largest operand / the output
        double e = 1;
        for (auto& l : P0.letters) if (t < P0.n_ops ? l.present[t] : l.in_out) e *= (double)l.extent;
        biggest = std::max(biggest, e);
    }
    if (P0.n_ops > ADB_RED_MAX_OPS) return true;                // the reduction kernel cannot take it at all
    if (naive * P0.n_ops < 4.0e6) return false;                 // one small launch beats a chain of launches
    Problem P = P0;
    double chain = 0;
    while (P.n_ops > 2) {
        const PairPlan pp = best_pair(P);
        chain += pp.iter;
        if (pp.inter > 4.0 * biggest && pp.inter * 8.0 > 256e6) return false;      // an outer-product blow-up: stay fused
        std::vector<Letter> keep;
        for (auto& l : P.letters) {
            const bool in_pair = l.present[pp.i] || l.present[pp.j];
            if (in_pair && !pair_letter_kept(P, l, pp.i, pp.j)) continue;
            Letter n = l;
            n.present[pp.i] = in_pair;
            for (int t = pp.j; t + 1 < P.n_ops; ++t) n.present[t] = n.present[t + 1];
            keep.push_back(n);
        }
        P.letters = keep;
        --P.n_ops;
    }
    double last = 1;
    for (auto& l : P.letters) last *= (double)l.extent;
    chain += last;
    return chain * 2.0 <= naive;
}

static int run_pairwise(const Problem& P0, cudaStream_t st, std::string* desc) {
    Problem P = P0;
    std::vector<double*> temps;
    std::string text = "pairwise[";
    int rc = 0;
    while (P.n_ops > 2 && rc == 0) {
        const PairPlan pp = best_pair(P);
        const int i = pp.i, j = pp.j;
        Problem S;
        S.n_ops = 2;
        S.op[0] = P.op[i]; S.op[1] = P.op[j];
        S.empty_out = S.empty_sum = false;
        std::vector<int> src;                                    // index into P.letters of every letter of S
        for (int k = 0; k < (int)P.letters.size(); ++k) {
            const Letter& l = P.letters[k];
            if (!l.present[i] && !l.present[j]) continue;
            Letter n;
            memset(&n, 0, sizeof n);
            n.c = l.c; n.extent = l.extent;
            n.s[0] = l.s[i]; n.present[0] = l.present[i];
            n.s[1] = l.s[j]; n.present[1] = l.present[j];
            n.in_out = pair_letter_kept(P, l, i, j);
            S.letters.push_back(n);
            src.push_back(k);
        }
        // dense layout of the intermediate: [both | first only | second only], each group in its source's stride order
        std::vector<int> order;
        for (int k = 0; k < (int)S.letters.size(); ++k) if (S.letters[k].in_out) order.push_back(k);
        if ((int)order.size() > ADB_MAX_RANK) { rc = set_error("einsum: intermediate of rank %d (max %d)", (int)order.size(), ADB_MAX_RANK); break; }
        auto grp = [&](int k) { const Letter& l = S.letters[k]; return l.present[0] && l.present[1] ? 0 : (l.present[0] ? 1 : 2); };
        auto key = [&](int k) { const Letter& l = S.letters[k]; const long long v = grp(k) == 2 ? l.s[1] : l.s[0]; return v < 0 ? -v : v; };
        std::sort(order.begin(), order.end(), [&](int a, int b) { return grp(a) != grp(b) ? grp(a) < grp(b) : key(a) > key(b); });
        long long elems = 1;
        for (int r = (int)order.size() - 1; r >= 0; --r) { S.letters[order[r]].os = elems; elems *= S.letters[order[r]].extent; }
        double* tmp = nullptr;
        if (!desc) {
            cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&tmp), (size_t)std::max(elems, 1LL) * sizeof(double), st);
            if (e != cudaSuccess) { rc = set_error("einsum: intermediate allocation failed: %s", cudaGetErrorString(e)); break; }
            temps.push_back(tmp);
        }
        S.out = tmp;
        S.out_t.ptr = tmp; S.out_t.rank = (int)order.size();
        for (int r = 0; r < (int)order.size(); ++r) { S.out_t.shape[r] = S.letters[order[r]].extent; S.out_t.stride[r] = S.letters[order[r]].os; }
        std::string sub;
        rc = run_problem(S, st, desc ? &sub : nullptr);
        if (desc) text += (text.size() > 9 ? " ; " : "") + sub;
        // the pair becomes one operand (slot i); slot j closes up
        std::vector<Letter> keep;
        for (int k = 0; k < (int)P.letters.size(); ++k) {
            Letter n = P.letters[k];
            const auto it = std::find(src.begin(), src.end(), k);
            if (it != src.end()) {
                const Letter& sl = S.letters[it - src.begin()];
                if (!sl.in_out) continue;                        // summed by this step
                n.s[i] = sl.os; n.present[i] = true;
            }
            for (int t = j; t + 1 < P.n_ops; ++t) { n.s[t] = n.s[t + 1]; n.present[t] = n.present[t + 1]; }
            n.s[P.n_ops - 1] = 0; n.present[P.n_ops - 1] = false;
            keep.push_back(n);
        }
        P.letters = keep;
        P.op[i] = tmp;
        for (int t = j; t + 1 < P.n_ops; ++t) P.op[t] = P.op[t + 1];
        --P.n_ops;
    }
    if (rc == 0) {
        std::string sub;
        rc = run_problem(P, st, desc ? &sub : nullptr);
        if (desc) *desc = text + " ; " + sub + "]";
    }
    for (double* t : temps) cudaFreeAsync(t, st);
    return rc;
}

static int run_problem(const Problem& P, cudaStream_t st, std::string* desc) {
    if (P.empty_out) { if (desc) *desc = "empty"; return 0; }
    if (P.empty_sum) return run_fill_zero(P, st, desc);
    bool has_sum = false;
    for (auto& l : P.letters) has_sum |= !l.in_out;
    if (!has_sum) return run_product(P, st, desc);
    if (P.n_ops >= 3 && pairwise_pays(P)) return run_pairwise(P, st, desc);
    adb_gemm_desc d;
    memset(&d, 0, sizeof d);
    bool swapped = false;
    if (try_gemm(P, d, &swapped)) {
        if (desc) {
            char b[256];
            snprintf(b, sizeof b, "gemm M=%lld N=%lld K=%lld batch=%lld a=(%lld,%lld,%lld) b=(%lld,%lld,%lld) c=(%lld,%lld,%lld)%s",
                     (long long)d.M, (long long)d.N, (long long)d.K, (long long)d.batch, (long long)d.a_sm, (long long)d.a_sk, (long long)d.a_sb,
                     (long long)d.b_sk, (long long)d.b_sn, (long long)d.b_sb, (long long)d.c_sm, (long long)d.c_sn, (long long)d.c_sb,
                     swapped ? " swapped" : "");
            *desc = b;
            long long r = 0;
            const int cut = ragged_edge(P, d, swapped, &r);
            if (cut >= 0) {
                snprintf(b, sizeof b, " + edge reduce of %lld along '%c'", r, P.letters[cut].c);
                *desc += b;
            }
            return 0;
        }
        // the 3xTF32 tcgen05 variant only pays off (and only has tiles) for reasonably large contractions
        if (g_precision == 1 && d.M >= 64 && d.N >= 64 && d.K >= 32) return gemm_tf32x3(&d, st);
        long long r = 0;
        const int cut = ragged_edge(P, d, swapped, &r);
        if (cut >= 0) {
            Problem main = P, edge = P;
            const Letter& l = P.letters[cut];
            main.letters[cut].extent = l.extent - r;
            adb_gemm_desc dm;
            memset(&dm, 0, sizeof dm);
            if (try_gemm(main, dm)) {
                for (int t = 0; t < P.n_ops; ++t)
                    if (l.present[t]) edge.op[t] = P.op[t] + (l.extent - r) * l.s[t];
                edge.out = P.out + (l.extent - r) * l.os;
                if (r == 1) edge.letters.erase(edge.letters.begin() + cut);      // (Problem keeps letters of extent > 1 only)
                else edge.letters[cut].extent = r;
                if (int rc = gemm_f64(&dm, st, 0)) return rc;
                return run_reduce(edge, st, nullptr);
            }
        }
        return gemm_f64(&d, st, 0);
    }
    return run_reduce(P, st, desc);
}

}  // namespace adb

extern "C" {

int adb_einsum(const char* op_str, int n_ops, const adb_tensor* ops, const adb_tensor* out, void* stream) {
    adb::Problem P;
    if (int rc = adb::parse(op_str, n_ops, ops, out, P)) return rc;
    return adb::run_problem(P, reinterpret_cast<cudaStream_t>(stream), nullptr);
}

int adb_einsum_describe(const char* op_str, int n_ops, const adb_tensor* ops, const adb_tensor* out, char* buf, size_t buflen) {
    adb::Problem P;
    if (int rc = adb::parse(op_str, n_ops, ops, out, P)) return rc;
    std::string desc;
    if (int rc = adb::run_problem(P, nullptr, &desc)) return rc;
    if (buf && buflen) {
        strncpy(buf, desc.c_str(), buflen - 1);
        buf[buflen - 1] = 0;
    }
    return 0;
}

// np.sum(x, axis=axes, keepdims=True) — reference reshape.py:14.  Reduced axes are those where out has extent 1.
int adb_reduce_sum(const adb_tensor* x, const adb_tensor* out, void* stream) {
    if (x->rank != out->rank) return adb::set_error("reduce_sum: rank mismatch %d vs %d", x->rank, out->rank);
    adb::Problem P;
    P.n_ops = 1;
    P.op[0] = x->ptr;
    P.out = out->ptr;
    P.out_t = *out;
    P.empty_out = P.empty_sum = false;
    for (int k = 0; k < x->rank; ++k) {
        const long long d = x->shape[k];
        const bool keep = out->shape[k] == d;
        if (!keep && out->shape[k] != 1) return adb::set_error("reduce_sum: out dim %d is %lld, expected %lld or 1", k, (long long)out->shape[k], d);
        if (d == 0) { (keep ? P.empty_out : P.empty_sum) = true; }
        if (d <= 1) continue;
        adb::Letter l;
        memset(&l, 0, sizeof l);
        l.c = (char)('a' + k);
        l.extent = d;
        l.in_out = keep;
        l.os = keep ? out->stride[k] : 0;
        l.s[0] = x->stride[k];
        l.present[0] = true;
        P.letters.push_back(l);
    }
    return adb::run_problem(P, reinterpret_cast<cudaStream_t>(stream), nullptr);
}

// np.sum(np.square(x)) — reference ops.py:369 (FrobeniusNorm); take_root: sqrt of it in the same launch.
static int sum_squares_impl(const adb_tensor* x, double* out_scalar, void* stream, int take_root) {
    adb::ReduceSpec S;
    memset(&S, 0, sizeof S);
    S.n_ops = 1;
    S.square = 1;
    S.post_sqrt = take_root;
    S.op[0] = x->ptr;
    S.out = out_scalar;
    // order dims by |stride| ascending and collapse
    std::vector<int> ax;
    for (int k = 0; k < x->rank; ++k) {
        if (x->shape[k] == 0) {
            adb_tensor o; o.ptr = out_scalar; o.rank = 0;
            adb_ew_instr code[2] = {{ADB_EW_LOAD | ADB_EW_SRC_IMM, 0, 0.0}, {ADB_EW_STORE_OUT, 0, 0.0}};
            return adb::ewise_run(code, 2, 0, nullptr, 1, &o, reinterpret_cast<cudaStream_t>(stream));
        }
        if (x->shape[k] > 1) ax.push_back(k);
    }
    std::sort(ax.begin(), ax.end(), [&](int a, int b) { return llabs(x->stride[a]) < llabs(x->stride[b]); });
    int n = 0;
    for (int k : ax) {
        if (n > 0 && x->stride[k] == S.op_rs[0][n - 1] * S.rdim[n - 1]) { S.rdim[n - 1] *= x->shape[k]; continue; }
        if (n == 4) return adb::set_error("sum_squares: more than 4 non-collapsible dims");
        S.rdim[n] = x->shape[k];
        S.op_rs[0][n] = x->stride[k];
        ++n;
    }
    S.n_r = n;
    S.n_o = 0;
    S.group = (n > 0 && S.rdim[0] >= 64) ? 32 : (n > 0 && S.rdim[0] >= 8 ? 8 : 1);
    return adb::reduce_run(S, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" {
int adb_sum_squares(const adb_tensor* x, double* out_scalar, void* stream) { return sum_squares_impl(x, out_scalar, stream, 0); }
int adb_norm2(const adb_tensor* x, double* out_scalar, void* stream) { return sum_squares_impl(x, out_scalar, stream, 1); }
}

}  // extern "C"
