// chol_plan.cu -- host side of the sparse multifrontal Cholesky: nested-dissection ordering on the node
// coordinates, elimination tree of fronts, symbolic factorisation (front boundaries), the index maps
// of the extend-add and the memory plan.  Pure host code (no CUDA call except the index upload).
//
// Ordering.  A node set is split at the median coordinate of one axis; the vertex separator is the
// smaller of the two one-sided boundary layers of the edge cut; all three axes are tried and the
// smallest separator wins.  Sets of at most `leaf_nodes` nodes become leaves.  Fronts of equal tree
// depth are independent and are processed as one batch ("level") by the numeric kernels.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <mutex>
#include <thread>

#include "chol.cuh"
#include "common.cuh"

namespace fers {
namespace {

// run fn(thread, begin, end) over [0, n) split into contiguous chunks, one std::thread per chunk
template <class F>
void parallel_chunks(int n, int max_threads, F fn) {
    const int t = std::max(1, std::min(max_threads, n / 8));
    if (t == 1) { fn(0, 0, n); return; }
    std::vector<std::thread> pool;
    for (int k = 0; k < t; ++k)
        pool.emplace_back([&, k]() { fn(k, (int)((long long)n * k / t), (int)((long long)n * (k + 1) / t)); });
    for (auto& th : pool) th.join();
}
constexpr int PLAN_THREADS = 8;

struct FrontBuild {
    int parent = -1;
    int depth = 0;
    unsigned long long key = 0;   // path from the root of the recursion (1 bit per split): orders siblings
    int key_bits = 0;
    std::vector<int> own;
    std::vector<int> children;
};

struct Dissector {
    int n;
    const int* rp;
    const int* ci;
    const double* xyz;
    int leaf;
    std::vector<uint8_t> part;
    std::vector<FrontBuild> fronts;
    std::mutex mu;                 // the two halves of the upper splits are dissected by different threads

    int new_front(int parent, std::vector<int>&& own, unsigned long long key, int key_bits) {
        FrontBuild f;
        f.parent = parent;
        f.own = std::move(own);
        f.key = key; f.key_bits = key_bits;
        std::lock_guard<std::mutex> lock(mu);
        f.depth = parent < 0 ? 0 : fronts[parent].depth + 1;
        fronts.push_back(std::move(f));
        const int id = (int)fronts.size() - 1;
        if (parent >= 0) fronts[parent].children.push_back(id);
        return id;
    }

    // split `nodes` along `axis` at the median; returns false when the axis cannot separate the set
    bool split_axis(const std::vector<int>& nodes, int axis, std::vector<int>& a, std::vector<int>& b) const {
        std::vector<double> c(nodes.size());
        for (size_t i = 0; i < nodes.size(); ++i) c[i] = xyz[(size_t)nodes[i] * 3 + axis];
        std::vector<double> tmp(c);
        const size_t mid = tmp.size() / 2;
        std::nth_element(tmp.begin(), tmp.begin() + mid, tmp.end());
        const double thr = tmp[mid];
        a.clear(); b.clear();
        for (size_t i = 0; i < nodes.size(); ++i) (c[i] < thr ? a : b).push_back(nodes[i]);
        if (a.empty()) {
            a.clear(); b.clear();
            for (size_t i = 0; i < nodes.size(); ++i) (c[i] <= thr ? a : b).push_back(nodes[i]);
        }
        return !a.empty() && !b.empty();
    }

    // one-sided boundary layers of the cut (part: 1 = a, 2 = b)
    // (part[] is shared between threads: a thread only writes the entries of its own node set, and the only
    // foreign entries it reads belong to separators of ancestors, which are final and 0)
    void separators(const std::vector<int>& a, const std::vector<int>& b, std::vector<int>& sa, std::vector<int>& sb) {
        separators(a, b, sa, sb, part);
    }
    void separators(const std::vector<int>& a, const std::vector<int>& b, std::vector<int>& sa, std::vector<int>& sb,
                    std::vector<uint8_t>& part) const {
        for (int v : a) part[v] = 1;
        for (int v : b) part[v] = 2;
        sa.clear(); sb.clear();
        for (int v : a)
            for (int k = rp[v]; k < rp[v + 1]; ++k)
                if (part[ci[k]] == 2) { sa.push_back(v); break; }
        for (int v : b)
            for (int k = rp[v]; k < rp[v + 1]; ++k)
                if (part[ci[k]] == 1) { sb.push_back(v); break; }
        for (int v : a) part[v] = 0;
        for (int v : b) part[v] = 0;
    }

    static constexpr int PAR_BITS = 4;            // splits whose halves go to separate threads (16-way); the three
                                                  // candidate axes of these splits are evaluated concurrently too
    static constexpr size_t PAR_MIN_NODES = 4096;

    void both(std::vector<int>&& a, std::vector<int>&& b, int parent, unsigned long long key, int bits) {
        if (bits < PAR_BITS && a.size() + b.size() >= PAR_MIN_NODES) {
            std::thread other([&, parent, key, bits]() { dissect(std::move(a), parent, key << 1, bits + 1); });
            dissect(std::move(b), parent, (key << 1) | 1ull, bits + 1);
            other.join();
        } else {
            dissect(std::move(a), parent, key << 1, bits + 1);
            dissect(std::move(b), parent, (key << 1) | 1ull, bits + 1);
        }
    }

    void dissect(std::vector<int>&& nodes, int parent, unsigned long long key = 1, int bits = 0) {
        if ((int)nodes.size() <= leaf || bits >= 62) {
            if (!nodes.empty()) new_front(parent, std::move(nodes), key, bits);
            return;
        }
        std::vector<int> best_a, best_b, best_sep;
        bool have = false, best_from_a = true;
        double best_score = 0.0;
        std::vector<int> a, b, sa, sb;
        // upper splits: one thread per candidate axis (own scratch marks); the choice below is taken in axis order
        // from the finished candidates, so the ordering does not depend on the threading
        const bool par_axes = bits < PAR_BITS && nodes.size() >= PAR_MIN_NODES;
        struct Candidate { bool ok = false; std::vector<int> a, b, sa, sb; };
        Candidate cand[3];
        if (par_axes) {
            auto eval = [&](int axis) {
                Candidate& c = cand[axis];
                std::vector<uint8_t> marks((size_t)n, 0);
                c.ok = split_axis(nodes, axis, c.a, c.b);
                if (c.ok) separators(c.a, c.b, c.sa, c.sb, marks);
            };
            std::thread t1(eval, 1), t2(eval, 2);
            eval(0);
            t1.join(); t2.join();
        }
        for (int axis = 0; axis < 3; ++axis) {
            if (par_axes) {
                if (!cand[axis].ok) continue;
                a.swap(cand[axis].a); b.swap(cand[axis].b); sa.swap(cand[axis].sa); sb.swap(cand[axis].sb);
            } else {
                if (!split_axis(nodes, axis, a, b)) continue;
                separators(a, b, sa, sb);
            }
            const bool from_a = sa.size() <= sb.size();
            const std::vector<int>& sep = from_a ? sa : sb;
            // balance guard: a median split is balanced unless many nodes share the median coordinate
            const size_t small = std::min(a.size(), b.size());
            const bool balanced = small * 5 >= nodes.size();
            const double score = (double)sep.size() * (balanced ? 1.0 : 4.0);
            if (!have || score < best_score) {
                have = true; best_from_a = from_a; best_score = score;
                best_a = a; best_b = b; best_sep = sep;
            }
        }
        if (!have) {   // coincident coordinates: split by position in the list
            const size_t mid = nodes.size() / 2;
            best_a.assign(nodes.begin(), nodes.begin() + mid);
            best_b.assign(nodes.begin() + mid, nodes.end());
            separators(best_a, best_b, sa, sb);
            best_from_a = sa.size() <= sb.size();
            best_sep = best_from_a ? sa : sb;
        }
        nodes.clear(); nodes.shrink_to_fit();
        if (best_sep.empty()) {   // the two halves are not connected: no separator front
            both(std::move(best_a), std::move(best_b), parent, key, bits);
            return;
        }
        // remove the separator from the side it was taken from
        for (int v : best_sep) part[v] = 3;
        std::vector<int>& side = best_from_a ? best_a : best_b;
        side.erase(std::remove_if(side.begin(), side.end(), [&](int v) { return part[v] == 3; }), side.end());
        for (int v : best_sep) part[v] = 0;
        const int f = new_front(parent, std::move(best_sep), key, bits);
        both(std::move(best_a), std::move(best_b), f, key, bits);
    }
};

}  // namespace

static int build_plan(int n_nodes, const int* rp, const int* ci, const double* xyz, int leaf_nodes, const uint8_t* drop,
                      CholPlanHost& P) {
    const bool timing = getenv("FERS_PLAN_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "chol_plan %-22s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    Dissector D;
    D.n = n_nodes; D.rp = rp; D.ci = ci; D.xyz = xyz; D.leaf = leaf_nodes < 1 ? 1 : leaf_nodes;
    D.part.assign(n_nodes, 0);
    {
        // nodes whose six DOFs are all restrained by supports take no part in the elimination
        std::vector<int> all;
        all.reserve(n_nodes);
        for (int v = 0; v < n_nodes; ++v)
            if (!drop || !drop[v]) all.push_back(v);
        if (all.empty()) { set_error("chol_plan: every node is fully restrained, nothing to factorise"); return 3; }
        P.n_active = (int)all.size();
        D.dissect(std::move(all), -1);
    }
    lap("dissection");
    std::vector<FrontBuild>& B = D.fronts;
    const int nf = (int)B.size();
    // front ids depend on thread timing; the recursion path does not: order siblings and roots by it
    auto path_less = [&](int x, int y) {
        const FrontBuild &fx = B[x], &fy = B[y];
        const int sx = 63 - fx.key_bits, sy = 63 - fy.key_bits;     // left-align the paths
        const unsigned long long kx = fx.key << sx, ky = fy.key << sy;
        return kx != ky ? kx < ky : fx.key_bits < fy.key_bits;
    };
    for (FrontBuild& fb : B) std::sort(fb.children.begin(), fb.children.end(), path_less);
    std::vector<int> roots;
    for (int r = 0; r < nf; ++r)
        if (B[r].parent < 0) roots.push_back(r);
    std::sort(roots.begin(), roots.end(), path_less);
    // post-order (children before parents), iterative
    std::vector<int> order; order.reserve(nf);
    {
        std::vector<std::pair<int, int>> stack;
        for (int r : roots) {
            stack.push_back({r, 0});
            while (!stack.empty()) {
                auto& top = stack.back();
                if (top.second < (int)B[top.first].children.size()) {
                    const int c = B[top.first].children[top.second++];
                    stack.push_back({c, 0});
                } else {
                    order.push_back(top.first);
                    stack.pop_back();
                }
            }
        }
    }
    std::vector<int> new_id(nf);
    for (int i = 0; i < nf; ++i) new_id[order[i]] = i;

    P.n_nodes = n_nodes;
    P.n_fronts = nf;
    P.perm.clear(); P.perm.reserve(n_nodes);
    P.pos.assign(n_nodes, -1);
    P.fronts.assign(nf, FrontDev{});
    P.front_first.assign(nf, 0);
    int max_depth = 0;
    for (int i = 0; i < nf; ++i) {
        FrontBuild& fb = B[order[i]];
        std::sort(fb.own.begin(), fb.own.end());
        P.front_first[i] = (int)P.perm.size();
        for (int v : fb.own) { P.pos[v] = (int)P.perm.size(); P.perm.push_back(v); }
        P.fronts[i].n1 = (int)fb.own.size();
        P.fronts[i].pad0 = P.front_first[i];
        P.fronts[i].parent = fb.parent < 0 ? -1 : new_id[fb.parent];
        max_depth = std::max(max_depth, fb.depth);
    }
    if ((int)P.perm.size() != P.n_active) { set_error("chol_plan: ordering lost nodes"); return 3; }

    lap("ordering");
    // symbolic factorisation: boundaries in post-order
    std::vector<std::vector<int>> bnd(nf);
    P.child.clear(); P.inv_off.clear();
    {
        // children before parents: fronts of equal depth are independent (deepest first), chunks of a depth on threads
        std::vector<std::vector<int>> by_depth(max_depth + 1);
        for (int i = 0; i < nf; ++i) by_depth[B[order[i]].depth].push_back(i);
        std::vector<std::vector<int>> stamps(PLAN_THREADS, std::vector<int>(n_nodes, -1));
        for (int d = max_depth; d >= 0; --d) {
            const std::vector<int>& ids = by_depth[d];
            parallel_chunks((int)ids.size(), PLAN_THREADS, [&](int t, int b0, int b1) {
                std::vector<int>& stamp = stamps[t];
                for (int q = b0; q < b1; ++q) {
                    const int i = ids[q];
                    const FrontBuild& fb = B[order[i]];
                    const int last = P.front_first[i] + P.fronts[i].n1;
                    std::vector<int>& out = bnd[i];
                    for (int v : fb.own)
                        for (int k = rp[v]; k < rp[v + 1]; ++k) {
                            const int w = ci[k];
                            if (P.pos[w] >= last && stamp[w] != i) { stamp[w] = i; out.push_back(w); }
                        }
                    for (int c : fb.children)
                        for (int w : bnd[new_id[c]])
                            if (P.pos[w] >= last && stamp[w] != i) { stamp[w] = i; out.push_back(w); }
                    std::sort(out.begin(), out.end(), [&](int x, int y) { return P.pos[x] < P.pos[y]; });
                    P.fronts[i].n2 = (int)out.size();
                }
            });
        }
    }

    lap("boundaries");
    // node lists, child lists and inverse maps
    long long factor = 0, rhs_rows = 0;
    P.flops = 0; P.nnz_l = 0; P.max_front_nodes = 0;
    std::vector<int> depth_of(nf);
    std::vector<long long> inv_base(nf, 0);
    {
        // offsets first (sequential, cheap), then the node lists and inverse maps of the fronts on threads
        long long n_nodes_list = 0, n_inv = 0;
        for (int i = 0; i < nf; ++i) {
            FrontDev& f = P.fronts[i];
            const FrontBuild& fb = B[order[i]];
            depth_of[i] = fb.depth;
            const int nt = f.n1 + f.n2;
            f.nodes_off = (int)n_nodes_list;
            n_nodes_list += nt;
            P.max_front_nodes = std::max(P.max_front_nodes, nt);
            f.child_begin = (int)P.child.size();
            inv_base[i] = n_inv;
            for (int c : fb.children) {
                P.child.push_back(new_id[c]);
                P.inv_off.push_back(n_inv);
                n_inv += nt;
            }
            f.child_end = (int)P.child.size();
            const long long m = 6LL * nt, k1 = 6LL * f.n1, m2 = 6LL * f.n2;
            f.panel_off = factor;
            factor += m * k1;
            f.dinv_off = factor;
            factor += ((k1 + CHOL_NB - 1) / CHOL_NB) * (long long)(CHOL_NB * CHOL_NB);
            f.rhs_off = rhs_rows;
            rhs_rows += m;
            P.flops += (double)k1 * k1 * k1 / 3.0 + (double)k1 * k1 * m2 + (double)k1 * m2 * m2;
            P.nnz_l += k1 * (k1 + 1) / 2 + m2 * k1;
        }
        if (n_nodes_list >= (1LL << 31)) { set_error("chol_plan: front node lists exceed 32-bit indexing"); return 3; }
        P.nodes.assign((size_t)n_nodes_list, 0);
        P.inv.assign((size_t)n_inv, -1);
        std::vector<int> failed(PLAN_THREADS, 0);
        std::vector<std::vector<int>> locs(PLAN_THREADS, std::vector<int>(n_nodes, -1));
        parallel_chunks(nf, PLAN_THREADS, [&](int t, int b0, int b1) {
            std::vector<int>& loc = locs[t];
            for (int i = b0; i < b1; ++i) {
                const FrontDev& f = P.fronts[i];
                const FrontBuild& fb = B[order[i]];
                int* nd = P.nodes.data() + f.nodes_off;
                int q = 0;
                for (int v : fb.own) nd[q++] = v;
                for (int v : bnd[i]) nd[q++] = v;
                const int nt = f.n1 + f.n2;
                for (int k = 0; k < nt; ++k) loc[nd[k]] = k;
                long long base = inv_base[i];
                for (int c : fb.children) {
                    const std::vector<int>& cb = bnd[new_id[c]];
                    for (int k = 0; k < (int)cb.size(); ++k) {
                        const int l = loc[cb[k]];
                        if (l < 0) { failed[t] = 1; continue; }
                        P.inv[(size_t)base + l] = k;
                    }
                    base += nt;
                }
                for (int k = 0; k < nt; ++k) loc[nd[k]] = -1;
            }
        });
        for (int f : failed)
            if (f) { set_error("chol_plan: child boundary node missing from parent front"); return 3; }
    }
    P.factor_doubles = factor;
    P.rhs_rows = rhs_rows;

    lap("node lists / inv maps");
    // levels: deepest depth first; the parent of a front of level l is in level l + 1
    P.n_levels = max_depth + 1;
    P.level_ptr.assign(P.n_levels + 1, 0);
    for (int i = 0; i < nf; ++i) P.level_ptr[max_depth - depth_of[i] + 1]++;
    for (int l = 0; l < P.n_levels; ++l) P.level_ptr[l + 1] += P.level_ptr[l];
    P.level_fronts.assign(nf, 0);
    {
        std::vector<int> cur(P.level_ptr.begin(), P.level_ptr.end() - 1);
        for (int i = 0; i < nf; ++i) P.level_fronts[cur[max_depth - depth_of[i]]++] = i;
    }
    P.level_max_n1.assign(P.n_levels, 0);
    P.level_max_n.assign(P.n_levels, 0);
    P.schur_doubles[0] = P.schur_doubles[1] = 0;
    for (int l = 0; l < P.n_levels; ++l) {
        long long s = 0;
        for (int q = P.level_ptr[l]; q < P.level_ptr[l + 1]; ++q) {
            FrontDev& f = P.fronts[P.level_fronts[q]];
            f.schur_off = s;
            s += 36LL * f.n2 * f.n2;
            P.level_max_n1[l] = std::max(P.level_max_n1[l], f.n1);
            P.level_max_n[l] = std::max(P.level_max_n[l], f.n1 + f.n2);
        }
        P.schur_doubles[l & 1] = std::max(P.schur_doubles[l & 1], s);
    }

    lap("levels");
    // K blocks of the lower triangle (in elimination order) -> panel positions, grouped by level
    std::vector<int> level_of(nf);
    for (int l = 0; l < P.n_levels; ++l)
        for (int q = P.level_ptr[l]; q < P.level_ptr[l + 1]; ++q) level_of[P.level_fronts[q]] = l;
    // per-front entry counts, then offsets (fronts of a level in index order), then the entries -- on threads
    std::vector<long long> front_cnt(nf, 0), front_off(nf, 0);
    parallel_chunks(nf, PLAN_THREADS, [&](int, int b0, int b1) {
        for (int i = b0; i < b1; ++i) {
            const FrontDev& f = P.fronts[i];
            long long n = 0;
            for (int q = 0; q < f.n1; ++q) {
                const int c = P.nodes[f.nodes_off + q];
                for (int k = rp[c]; k < rp[c + 1]; ++k)
                    if (P.pos[ci[k]] >= P.pos[c]) ++n;
            }
            front_cnt[i] = n;
        }
    });
    std::vector<long long> cnt(P.n_levels + 1, 0);
    for (int i = 0; i < nf; ++i) cnt[level_of[i] + 1] += front_cnt[i];
    for (int l = 0; l < P.n_levels; ++l) cnt[l + 1] += cnt[l];
    {
        std::vector<long long> cur(cnt.begin(), cnt.end() - 1);
        for (int i = 0; i < nf; ++i) { front_off[i] = cur[level_of[i]]; cur[level_of[i]] += front_cnt[i]; }
    }
    P.kscatter_ptr = cnt;
    P.kscatter.assign((size_t)cnt[P.n_levels], KScatter{});
    {
        std::vector<int> failed(PLAN_THREADS, 0);
        std::vector<std::vector<int>> locs(PLAN_THREADS, std::vector<int>(n_nodes, -1));
        // local index of a node inside a front: own nodes by position, boundary nodes through the front's node list
        parallel_chunks(nf, PLAN_THREADS, [&](int t, int b0, int b1) {
            std::vector<int>& loc = locs[t];
            for (int i = b0; i < b1; ++i) {
                const FrontDev& f = P.fronts[i];
                const int nt = f.n1 + f.n2;
                size_t at = (size_t)front_off[i];
                for (int q = 0; q < nt; ++q) loc[P.nodes[f.nodes_off + q]] = q;
                for (int q = 0; q < f.n1; ++q) {
                    const int c = P.nodes[f.nodes_off + q];          // column node (own)
                    for (int k = rp[c]; k < rp[c + 1]; ++k) {
                        const int r = ci[k];                          // row node; block (r, c) = transpose slot of (c, r)
                        if (P.pos[r] < P.pos[c]) continue;
                        // we need the block stored at row r, column c: find it in row r
                        const int* b = std::lower_bound(ci + rp[r], ci + rp[r + 1], c);
                        if (loc[r] < 0) { failed[t] |= 1; continue; }
                        if (b == ci + rp[r + 1] || *b != c) { failed[t] |= 2; continue; }
                        KScatter e; e.front = i; e.k = (int)(b - ci); e.bi = loc[r]; e.bj = q;
                        P.kscatter[at++] = e;
                    }
                }
                for (int q = 0; q < nt; ++q) loc[P.nodes[f.nodes_off + q]] = -1;
            }
        });
        for (int f : failed) {
            if (f & 1) { set_error("chol_plan: matrix entry outside its front"); return 3; }
            if (f & 2) { set_error("chol_plan: pattern is not structurally symmetric"); return 3; }
        }
    }

    lap("k scatter list");
    // device index buffer layout
    size_t o = 0;
    auto place = [&](size_t bytes) { const size_t at = o; o = align_up(o + bytes); return at; };
    P.off_fronts = place(sizeof(FrontDev) * (size_t)nf);
    P.off_nodes = place(sizeof(int) * P.nodes.size());
    P.off_child = place(sizeof(int) * P.child.size());
    P.off_inv_off = place(sizeof(long long) * P.inv_off.size());
    P.off_inv = place(sizeof(int) * P.inv.size());
    P.off_level_fronts = place(sizeof(int) * P.level_fronts.size());
    P.off_kscatter = place(sizeof(KScatter) * P.kscatter.size());
    P.off_pos = place(sizeof(int) * P.pos.size());
    P.index_bytes = o;
    return 0;
}

}  // namespace fers

using namespace fers;

extern "C" {

int fers_b200_chol_plan_create(int n_nodes, const int* rowptr, const int* colidx, const double* xyz, int leaf_nodes,
                               const uint8_t* node_restrained, void** plan_out) {
    FERS_REQUIRE(n_nodes > 0 && rowptr && colidx && xyz && plan_out, "chol_plan_create: bad arguments");
    CholPlanHost* P = new CholPlanHost();
    const int rc = build_plan(n_nodes, rowptr, colidx, xyz, leaf_nodes, node_restrained, *P);
    if (rc) { delete P; return rc; }
    *plan_out = P;
    return 0;
}

void fers_b200_chol_plan_destroy(void* plan) { delete static_cast<CholPlanHost*>(plan); }

/* [host] the block-CSR pattern fers_b200_pattern_build produces on the device (sorted unique (row, col) pairs of every
 * diagonal, both orientations of every node pair, the six ordered pairs of every TRI3 plate), built on the host so
 * that the factorisation plan can start before the device pattern has been copied back. */
int fers_b200_host_pattern(int n_nodes, int n_pairs, const int* p1, const int* p2, int n_plates, const int* p_nodes,
                           int* rowptr, int* colidx, int cap, int* nnzb_out) {
    FERS_REQUIRE(n_nodes > 0 && rowptr && colidx && nnzb_out, "host_pattern: bad arguments");
    FERS_REQUIRE(n_pairs == 0 || (p1 && p2), "host_pattern: node pairs missing");
    FERS_REQUIRE(n_plates == 0 || p_nodes, "host_pattern: plate nodes missing");
    std::vector<int> cnt((size_t)n_nodes + 1, 0);
    auto ok = [&](int v) { return v >= 0 && v < n_nodes; };
    for (int i = 0; i < n_pairs; ++i) {
        FERS_REQUIRE(ok(p1[i]) && ok(p2[i]), "host_pattern: node index out of range");
        ++cnt[p1[i] + 1]; ++cnt[p2[i] + 1];
    }
    for (int e = 0; e < 3 * n_plates; ++e) {
        FERS_REQUIRE(ok(p_nodes[e]), "host_pattern: plate node index out of range");
        cnt[p_nodes[e] + 1] += 2;
    }
    for (int i = 0; i < n_nodes; ++i) cnt[i + 1] += cnt[i] + 1;          // + the diagonal
    std::vector<int> raw((size_t)cnt[n_nodes]);
    std::vector<int> fill(cnt.begin(), cnt.end() - 1);
    for (int i = 0; i < n_nodes; ++i) raw[fill[i]++] = i;
    for (int i = 0; i < n_pairs; ++i) { raw[fill[p1[i]]++] = p2[i]; raw[fill[p2[i]]++] = p1[i]; }
    for (int e = 0; e < n_plates; ++e)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                if (a != b) raw[fill[p_nodes[3 * e + a]]++] = p_nodes[3 * e + b];
    // sort + unique per row (rows are short), in parallel chunks; compact afterwards
    std::vector<int> len((size_t)n_nodes);
    parallel_chunks(n_nodes, PLAN_THREADS, [&](int, int b0, int b1) {
        for (int i = b0; i < b1; ++i) {
            int* lo = raw.data() + cnt[i];
            int* hi = raw.data() + cnt[i + 1];
            std::sort(lo, hi);
            len[i] = (int)(std::unique(lo, hi) - lo);
        }
    });
    long long nnzb = 0;
    for (int i = 0; i < n_nodes; ++i) nnzb += len[i];
    FERS_REQUIRE(nnzb <= cap, "host_pattern: colidx capacity too small");
    int o = 0;
    for (int i = 0; i < n_nodes; ++i) {
        rowptr[i] = o;
        std::copy(raw.data() + cnt[i], raw.data() + cnt[i] + len[i], colidx + o);
        o += len[i];
    }
    rowptr[n_nodes] = o;
    *nnzb_out = o;
    return 0;
}

int fers_b200_chol_plan_info(void* plan, long long* out, int n_out) {
    FERS_REQUIRE(plan && out && n_out >= 12, "chol_plan_info: need 12 outputs");
    const CholPlanHost& P = *static_cast<CholPlanHost*>(plan);
    out[0] = P.n_fronts;
    out[1] = P.n_levels;
    out[2] = P.max_front_nodes;
    out[3] = P.factor_doubles;
    out[4] = P.schur_doubles[0];
    out[5] = P.schur_doubles[1];
    out[6] = P.rhs_rows;
    out[7] = (long long)P.index_bytes;
    out[8] = (long long)(P.flops / 1e6);   // Mflop
    out[9] = P.nnz_l;
    out[10] = (long long)P.kscatter.size();
    out[11] = P.n_levels ? P.fronts[P.level_fronts[P.level_ptr[P.n_levels - 1]]].n1 : 0;   // root separator nodes
    return 0;
}

int fers_b200_chol_plan_perm(void* plan, int* perm_out) {
    FERS_REQUIRE(plan && perm_out, "chol_plan_perm: bad arguments");
    const CholPlanHost& P = *static_cast<CholPlanHost*>(plan);
    std::fill(perm_out, perm_out + P.n_nodes, -1);          // fully restrained nodes are not eliminated
    std::copy(P.perm.begin(), P.perm.end(), perm_out);
    return 0;
}

/* per-level statistics: out[level*4 + {0,1,2,3}] = fronts, max own nodes, max front nodes, Mflop */
int fers_b200_chol_plan_levels(void* plan, long long* out, int n_levels_cap) {
    FERS_REQUIRE(plan && out, "chol_plan_levels: bad arguments");
    const CholPlanHost& P = *static_cast<CholPlanHost*>(plan);
    FERS_REQUIRE(n_levels_cap >= P.n_levels, "chol_plan_levels: output too small");
    for (int l = 0; l < P.n_levels; ++l) {
        double fl = 0;
        for (int q = P.level_ptr[l]; q < P.level_ptr[l + 1]; ++q) {
            const FrontDev& f = P.fronts[P.level_fronts[q]];
            const double k1 = 6.0 * f.n1, m2 = 6.0 * f.n2;
            fl += k1 * k1 * k1 / 3.0 + k1 * k1 * m2 + k1 * m2 * m2;
        }
        out[l * 4 + 0] = P.level_ptr[l + 1] - P.level_ptr[l];
        out[l * 4 + 1] = P.level_max_n1[l];
        out[l * 4 + 2] = P.level_max_n[l];
        out[l * 4 + 3] = (long long)(fl / 1e6);
    }
    return 0;
}

/* Tree-parallel factorisation on 2^k GPUs (strong scaling of ONE model).  The top k levels of the front tree must be a
 * complete binary tree; its fronts are numbered (depth d, index j) from the root (0, 0) by the order of the child lists.
 * Rank s keeps the subtree under front (k, s); an upper front (d, j) belongs to the leftmost rank below it, j << (k - d).
 * Everything else is dropped from this rank's level lists and from its scatter list of the K blocks; offsets into the
 * factor store, the Schur arenas and the right-hand-side stores stay those of the whole plan, so what one rank computes can
 * be copied to the same place on another.  Call before chol_upload.
 * out: [0] k, [1] level of the root, [2] fronts kept, [3] K blocks kept, then the 2^(k+1) - 1 upper fronts depth by depth,
 * then the first elimination position of each of the 2^k subtrees (its last is that of its top front). */
int fers_b200_chol_plan_partition(void* plan, int rank, int world, long long* out, int n_out) {
    FERS_REQUIRE(plan && out, "chol_plan_partition: bad arguments");
    int k = 0;
    while ((1 << k) < world) ++k;
    FERS_REQUIRE(world >= 2 && world <= 64 && (1 << k) == world && rank >= 0 && rank < world,
                 "chol_plan_partition: the number of ranks must be a power of two (2 .. 64)");
    const int n_top = (2 << k) - 1;
    FERS_REQUIRE(n_out >= 4 + n_top + world, "chol_plan_partition: output too small");
    CholPlanHost& P = *static_cast<CholPlanHost*>(plan);
    FERS_REQUIRE(P.n_levels > k, "chol_plan_partition: the front tree is shallower than the rank tree");
    const int ls = P.n_levels - 1;
    FERS_REQUIRE(P.level_ptr[ls + 1] - P.level_ptr[ls] == 1, "chol_plan_partition: the top level holds more than one front");
    std::vector<int> top(n_top, -1), owner(P.n_fronts, -1);
    top[0] = P.level_fronts[P.level_ptr[ls]];
    for (int d = 0; d < k; ++d)
        for (int j = 0; j < (1 << d); ++j) {
            const FrontDev& F = P.fronts[top[(1 << d) - 1 + j]];
            FERS_REQUIRE(F.child_end - F.child_begin == 2, "chol_plan_partition: an upper front does not have exactly two children");
            top[(2 << d) - 1 + 2 * j] = P.child[F.child_begin];
            top[(2 << d) - 1 + 2 * j + 1] = P.child[F.child_begin + 1];
        }
    for (int d = 0; d <= k; ++d)
        for (int j = 0; j < (1 << d); ++j) owner[top[(1 << d) - 1 + j]] = j << (k - d);
    // every other front belongs to the rank of the depth-k front above it (fronts are in post-order: the climb from an
    // unresolved front passes unresolved fronts only, until it meets an upper front)
    std::vector<char> is_leaf_top(P.n_fronts, 0);
    for (int j = 0; j < world; ++j) is_leaf_top[top[world - 1 + j]] = 1;
    std::vector<int> under(P.n_fronts, -1);
    for (int f = 0; f < P.n_fronts; ++f) {
        if (owner[f] >= 0) continue;
        int a = f, hops = 0;
        while (owner[a] < 0 && P.fronts[a].parent >= 0 && hops++ < P.n_fronts) a = P.fronts[a].parent;
        FERS_REQUIRE(owner[a] >= 0 && is_leaf_top[a], "chol_plan_partition: a front does not hang under one of the 2^k subtree tops");
        under[f] = owner[a];
    }
    for (int f = 0; f < P.n_fronts; ++f)
        if (owner[f] < 0) owner[f] = under[f];
    std::vector<long long> lo(world, 1LL << 60), hi(world, 0), cnt(world, 0);
    std::vector<char> is_top(P.n_fronts, 0);
    for (int t = 0; t < world - 1; ++t) is_top[top[t]] = 1;          // depths 0 .. k-1
    for (int f = 0; f < P.n_fronts; ++f) {
        if (is_top[f]) continue;
        const int o = owner[f];
        lo[o] = std::min<long long>(lo[o], P.fronts[f].pad0);
        hi[o] = std::max<long long>(hi[o], (long long)P.fronts[f].pad0 + P.fronts[f].n1);
        cnt[o] += P.fronts[f].n1;
    }
    for (int o = 0; o < world; ++o) {
        FERS_REQUIRE(hi[o] - lo[o] == cnt[o], "chol_plan_partition: a subtree is not contiguous in the elimination order");
        const FrontDev& T = P.fronts[top[world - 1 + o]];
        FERS_REQUIRE(hi[o] == (long long)T.pad0 + T.n1, "chol_plan_partition: a subtree does not end with its top front");
    }
    std::vector<int> lf, lp(P.n_levels + 1, 0);
    std::vector<KScatter> ks;
    std::vector<long long> kp(P.n_levels + 1, 0);
    for (int l = 0; l < P.n_levels; ++l) {
        int mx1 = 0, mx = 0;
        for (int q = P.level_ptr[l]; q < P.level_ptr[l + 1]; ++q) {
            const int f = P.level_fronts[q];
            if (owner[f] != rank) continue;
            lf.push_back(f);
            mx1 = std::max(mx1, P.fronts[f].n1);
            mx = std::max(mx, P.fronts[f].n1 + P.fronts[f].n2);
        }
        lp[l + 1] = (int)lf.size();
        P.level_max_n1[l] = mx1; P.level_max_n[l] = mx;
        for (long long q = P.kscatter_ptr[l]; q < P.kscatter_ptr[l + 1]; ++q)
            if (owner[P.kscatter[q].front] == rank) ks.push_back(P.kscatter[q]);
        kp[l + 1] = (long long)ks.size();
    }
    P.level_fronts.swap(lf); P.level_ptr.swap(lp); P.kscatter.swap(ks); P.kscatter_ptr.swap(kp);
    out[0] = k; out[1] = ls; out[2] = (long long)P.level_fronts.size(); out[3] = (long long)P.kscatter.size();
    for (int t = 0; t < n_top; ++t) out[4 + t] = top[t];
    for (int o = 0; o < world; ++o) out[4 + n_top + o] = lo[o];
    return 0;
}

}  // extern "C"

namespace fers {
// serialise the index arrays into one buffer laid out as CholPlanHost::off_* describe
void chol_plan_serialise(const CholPlanHost& P, unsigned char* out) {
    memcpy(out + P.off_fronts, P.fronts.data(), sizeof(FrontDev) * P.fronts.size());
    memcpy(out + P.off_nodes, P.nodes.data(), sizeof(int) * P.nodes.size());
    memcpy(out + P.off_child, P.child.data(), sizeof(int) * P.child.size());
    memcpy(out + P.off_inv_off, P.inv_off.data(), sizeof(long long) * P.inv_off.size());
    memcpy(out + P.off_inv, P.inv.data(), sizeof(int) * P.inv.size());
    memcpy(out + P.off_level_fronts, P.level_fronts.data(), sizeof(int) * P.level_fronts.size());
    memcpy(out + P.off_kscatter, P.kscatter.data(), sizeof(KScatter) * P.kscatter.size());
    memcpy(out + P.off_pos, P.pos.data(), sizeof(int) * P.pos.size());
}
}  // namespace fers

extern "C" {

/* [host] copy of the index buffer (index_bytes of chol_plan_info) and of its section offsets:
 * offsets_out[8] = fronts, nodes, child, inv_off, inv, level_fronts, kscatter, n_child_entries;
 * level_ptr_out[n_levels+1], kscatter_ptr_out[n_levels+1]. */
int fers_b200_chol_plan_export(void* plan, void* index_out, long long* offsets_out, int* level_ptr_out,
                               long long* kscatter_ptr_out) {
    FERS_REQUIRE(plan && index_out && offsets_out && level_ptr_out && kscatter_ptr_out, "chol_plan_export: bad arguments");
    const CholPlanHost& P = *static_cast<CholPlanHost*>(plan);
    chol_plan_serialise(P, static_cast<unsigned char*>(index_out));
    offsets_out[0] = (long long)P.off_fronts; offsets_out[1] = (long long)P.off_nodes; offsets_out[2] = (long long)P.off_child;
    offsets_out[3] = (long long)P.off_inv_off; offsets_out[4] = (long long)P.off_inv;
    offsets_out[5] = (long long)P.off_level_fronts; offsets_out[6] = (long long)P.off_kscatter;
    offsets_out[7] = (long long)P.child.size();
    std::copy(P.level_ptr.begin(), P.level_ptr.end(), level_ptr_out);
    std::copy(P.kscatter_ptr.begin(), P.kscatter_ptr.end(), kscatter_ptr_out);
    return 0;
}

}  // extern "C"
