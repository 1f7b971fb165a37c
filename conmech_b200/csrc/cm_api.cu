// C-ABI entry points of libconmech_b200 (see include/conmech_b200.h) and the host-side,
// once-per-model precomputation: node->element incidence, the static solver ordering
// (reverse Cuthill-McKee on the node graph), envelope statistics and the workspace layout.
#include "cm_common.cuh"
#include <chrono>
#include <thread>
#include <cstdlib>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <queue>

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t _e = (call);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return fail(CM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));         \
    } while (0)

#define CM_STAGE_SLOTS 6
struct cm_handle {
    int device = 0;
    cm_model_dev m{};
    cm_model_info info{};
    cm_ws_layout L{};
    std::vector<void *> allocs;          // device allocations owned by the handle
    std::vector<int> order;              // host copy of the ordering
    // geometry tables (stacked over the geometry variants; replaced as a set by cm_set_geometries)
    double *d_xyz = nullptr, *d_Kg = nullptr, *d_Kt = nullptr, *d_R3 = nullptr, *d_elen = nullptr;
    std::vector<int32_t> conn;           // host copy of the connectivity (length checks of new geometries)
    int *d_geom = nullptr;               // staging of the per-subset geometry indices of the host entry point
    size_t geom_cap = 0;
    long long *d_rowoff = nullptr;       // device copy of the caller's eR_row_offsets (compact element rows)
    size_t rowoff_cap = 0;
    size_t geom_tables_cap = 1;          // geometry variants the stacked tables have room for
    double k1_ms = 0.0;                  // device time of the most recent element-table launch
    // load tables
    double *d_q = nullptr, *d_pnodal = nullptr;
    double rot_tol = 0.0;
    int variant = 1;                     // 1 = FP64 tensor-core dataflow kernel, CTA size chosen per launch (default); 2/3/6 = the same with 4/8/16-warp CTAs;
                                         // 4/5 = 4/8-warp CTAs instrumented; 0 = scalar validation
    // counters / timing
    long long launches = 0;
    int timing = 0;
    int last_waves = 0;                  // waves of the most recent cm_solve_many call
    double stage_ms[3] = {0, 0, 0};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // staging for the host-buffer entry point (grow only)
    char *stage = nullptr;
    size_t stage_bytes = 0;
    char *wsbuf = nullptr;
    size_t ws_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t aux_stream = nullptr;   // second stream of a prefix-sharing wave (assembly of the tails beside the trunks' factorisation)
    cudaEvent_t aux_ev[3] = {nullptr, nullptr, nullptr};
    // internal lanes: consecutive waves of one call rotate over them (each with its own share
    // of the workspace), so that the drain of one wave's factorisation overlaps the next wave
    cudaStream_t lane_stream[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t lane_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};     // fork, join per lane
    // host-buffer entry point: ring of device staging slots for the outputs of one wave each, drained
    // by a copy stream while the lanes compute the following waves
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t slot_done[CM_STAGE_SLOTS] = {}, slot_free[CM_STAGE_SLOTS] = {};
    cudaEvent_t copy_ev = nullptr;
    // one copy stream per lane: the results of a wave leave when ITS kernels are done, not behind the copies of waves
    // that were enqueued earlier but finish later on another lane
    cudaStream_t copy_lane[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t copy_lane_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int overlap = 1;
    int n_lanes = 3;                     // CM_LANES=1..4: 2 lanes 250 k, 3 lanes 255 k, 4 lanes 255 k solves/s on config 3
};

template <typename T>
static int dev_upload(cm_handle *h, const T *src, size_t n, T **dst) {
    void *p = nullptr;
    CU(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    h->allocs.push_back(p);
    if (n && src) CU(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    *dst = static_cast<T *>(p);
    return CM_OK;
}

// ---- reverse Cuthill-McKee over the node graph (all components), host side
static std::vector<int> rcm_order(int nV, const std::vector<std::vector<int>> &adj) {
    std::vector<int> order;
    order.reserve(nV);
    std::vector<char> seen(nV, 0);
    std::vector<int> level(nV);
    auto bfs_far = [&](int root, std::vector<int> &comp) {
        // returns a node of the last BFS level with minimum degree, fills comp with the component
        comp.clear();
        std::queue<int> q;
        std::vector<int> touched;
        level[root] = 0;
        q.push(root);
        std::vector<char> mark(nV, 0);
        mark[root] = 1;
        int last = 0;
        while (!q.empty()) {
            int v = q.front(); q.pop();
            comp.push_back(v);
            last = level[v];
            for (int w : adj[v]) if (!mark[w]) { mark[w] = 1; level[w] = level[v] + 1; q.push(w); }
        }
        int best = root;
        size_t bestdeg = (size_t)-1;
        for (int v : comp) if (level[v] == last && adj[v].size() < bestdeg) { best = v; bestdeg = adj[v].size(); }
        return std::make_pair(best, last);
    };
    for (int s = 0; s < nV; ++s) {
        if (seen[s]) continue;
        std::vector<int> comp;
        // pseudo-peripheral start
        int root = s, ecc = -1;
        for (int it = 0; it < 8; ++it) {
            auto r = bfs_far(root, comp);
            if (r.second <= ecc) break;
            ecc = r.second;
            root = r.first;
        }
        // Cuthill-McKee from root
        std::vector<int> cm;
        std::queue<int> q;
        seen[root] = 1;
        q.push(root);
        while (!q.empty()) {
            int v = q.front(); q.pop();
            cm.push_back(v);
            std::vector<int> nb;
            for (int w : adj[v]) if (!seen[w]) { seen[w] = 1; nb.push_back(w); }
            std::sort(nb.begin(), nb.end(), [&](int a, int b) {
                return adj[a].size() != adj[b].size() ? adj[a].size() < adj[b].size() : a < b;
            });
            for (int w : nb) q.push(w);
        }
        order.insert(order.end(), cm.rbegin(), cm.rend());
    }
    return order;
}

// envelope statistics of the full structure for an ordering: half bandwidth, sum h^2, n_free
static void envelope_stats(int nV, int nE, const int32_t *conn, const std::vector<int> &nfree,
                           const std::vector<int> &order, int *n_free, int *halfbw, long long *flops) {
    std::vector<int> base(nV, 0);
    int run = 0;
    for (int p = 0; p < nV; ++p) { base[order[p]] = run; run += nfree[order[p]]; }
    *n_free = run;
    std::vector<int> first(nV);
    for (int v = 0; v < nV; ++v) first[v] = base[v];
    for (int e = 0; e < nE; ++e) {
        int u = conn[2 * e], v = conn[2 * e + 1];
        if (nfree[u] == 0 || nfree[v] == 0) continue;
        first[u] = std::min(first[u], base[v]);
        first[v] = std::min(first[v], base[u]);
    }
    int bw = 0;
    long long fl = 0;
    for (int v = 0; v < nV; ++v)
        for (int i = 0; i < nfree[v]; ++i) {
            int r = base[v] + i;
            int hgt = r - first[v];
            bw = std::max(bw, hgt);
            fl += (long long)hgt * hgt;
        }
    *halfbw = bw;
    *flops = fl;
}

extern "C" {

const char *cm_last_error(void) { return g_err.c_str(); }
const char *cm_version(void) { return "conmech_b200 0.1.0 (sm_100a)"; }

int cm_create(const cm_model_desc *d, cm_handle **out) {
    if (!d || !out) return fail(CM_ERR_INVALID, "cm_create: null argument");
    if (d->n_nodes <= 0 || d->n_elems <= 0) return fail(CM_ERR_INVALID, "cm_create: empty model");
    if (d->n_supports <= 0)
        return fail(CM_ERR_INVALID, "there needs to be at least one support (fixed) vertex in the model!");
    if (!d->xyz || !d->conn || !d->props || !d->cr || !d->sup_node || !d->sup_cond)
        return fail(CM_ERR_INVALID, "cm_create: null table in the model description");
    const int nV = d->n_nodes, nE = d->n_elems, nS = d->n_supports;
    {   // supports: in range, one entry per node (checked before any CUDA call, like the element checks below)
        std::vector<char> is_sup(nV, 0);
        for (int s = 0; s < nS; ++s) {
            const int v = d->sup_node[s];
            if (v < 0 || v >= nV) return fail(CM_ERR_INVALID, "cm_create: support node out of range");
            if (is_sup[v]) return fail(CM_ERR_INVALID, "cm_create: duplicate support node");
            is_sup[v] = 1;
        }
        if (d->node_order) {
            std::vector<char> seen(nV, 0);
            for (int p = 0; p < nV; ++p) {
                const int v = d->node_order[p];
                if (v < 0 || v >= nV || seen[v]) return fail(CM_ERR_INVALID, "cm_create: node_order is not a permutation of the node ids");
                seen[v] = 1;
            }
        }
    }
    for (int e = 0; e < nE; ++e) {
        int u = d->conn[2 * e], v = d->conn[2 * e + 1];
        if (u < 0 || u >= nV || v < 0 || v >= nV || u == v)
            return fail(CM_ERR_INVALID, "cm_create: element " + std::to_string(e) + " has invalid end nodes");
        double dx = d->xyz[3 * u] - d->xyz[3 * v], dy = d->xyz[3 * u + 1] - d->xyz[3 * v + 1],
               dz = d->xyz[3 * u + 2] - d->xyz[3 * v + 2];
        if (std::sqrt(dx * dx + dy * dy + dz * dz) < CM_EPS)      // numpy_stiffness_assembly.py:288-289
            return fail(CM_ERR_INVALID, "E#" + std::to_string(e) + " - Bar length too small!");
        if (!(std::fabs(d->props[7 * e + 4]) > CM_EPS) || !(std::fabs(d->props[7 * e + 3]) > CM_EPS))
            return fail(CM_ERR_INVALID, "E#" + std::to_string(e) + " - E and Iz must be non-zero");   // :146
    }
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (d->device < 0 || d->device >= ndev) return fail(CM_ERR_INVALID, "cm_create: no such CUDA device");
    CU(cudaSetDevice(d->device));

    cm_handle *h = new cm_handle();
    h->device = d->device;
    cm_model_dev &m = h->m;
    m.device = d->device;
    m.n_sm = 0;
    cudaDeviceGetAttribute(&m.n_sm, cudaDevAttrMultiProcessorCount, d->device);
    if (m.n_sm <= 0) m.n_sm = 148;
    m.nV = nV; m.nE = nE; m.nS = nS; m.n_lc = 0;
    m.n_geom = 1;
    h->conn.assign(d->conn, d->conn + 2 * (size_t)nE);
    m.mask_words = (nE + 31) / 32;
    m.trans_tol = 1e-3;
    h->rot_tol = 3.14159265358979323846 / 180 * 3;

    // incidence (ascending element ids per node) and node adjacency
    std::vector<int> ne_ptr(nV + 1, 0);
    for (int e = 0; e < nE; ++e) { ne_ptr[d->conn[2 * e] + 1]++; ne_ptr[d->conn[2 * e + 1] + 1]++; }
    for (int v = 0; v < nV; ++v) ne_ptr[v + 1] += ne_ptr[v];
    std::vector<int> ne_elem(2 * (size_t)nE), fill(ne_ptr.begin(), ne_ptr.end() - 1);
    std::vector<unsigned char> ne_end(2 * (size_t)nE);
    for (int e = 0; e < nE; ++e)
        for (int k = 0; k < 2; ++k) {
            int v = d->conn[2 * e + k];
            ne_elem[fill[v]] = e;
            ne_end[fill[v]] = (unsigned char)k;
            fill[v]++;
        }
    std::vector<std::vector<int>> adj(nV);
    for (int e = 0; e < nE; ++e) {
        adj[d->conn[2 * e]].push_back(d->conn[2 * e + 1]);
        adj[d->conn[2 * e + 1]].push_back(d->conn[2 * e]);
    }
    for (auto &a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }

    std::vector<int> node_sup(nV, -1), nfree(nV, 6);
    for (int s = 0; s < nS; ++s) {
        int v = d->sup_node[s];
        if (v < 0 || v >= nV) { delete h; return fail(CM_ERR_INVALID, "cm_create: support node out of range"); }
        if (node_sup[v] >= 0) { delete h; return fail(CM_ERR_INVALID, "cm_create: duplicate support node"); }
        node_sup[v] = s;
        int nf = 6;
        for (int i = 0; i < 6; ++i) nf -= d->sup_cond[6 * s + i] ? 1 : 0;
        nfree[v] = nf;
    }

    // static ordering: best of natural and RCM by profile flops of the full structure
    std::vector<int> nat(nV);
    std::iota(nat.begin(), nat.end(), 0);
    std::vector<int> rcm = rcm_order(nV, adj);
    int nf0, bw0, nf1, bw1;
    long long fl0, fl1;
    envelope_stats(nV, nE, d->conn, nfree, nat, &nf0, &bw0, &fl0);
    envelope_stats(nV, nE, d->conn, nfree, rcm, &nf1, &bw1, &fl1);
    bool use_rcm = fl1 < fl0;
    h->order = use_rcm ? rcm : nat;
    int n_free_full = nf0, halfbw = use_rcm ? bw1 : bw0;
    if (d->node_order) {                 // the caller's ordering (construction sequences: order of first appearance)
        std::vector<int> ord(d->node_order, d->node_order + nV);
        std::vector<char> seen(nV, 0);
        for (int p = 0; p < nV; ++p) {
            if (ord[p] < 0 || ord[p] >= nV || seen[ord[p]]) { delete h; return fail(CM_ERR_INVALID, "cm_create: node_order is not a permutation of the node ids"); }
            seen[ord[p]] = 1;
        }
        envelope_stats(nV, nE, d->conn, nfree, ord, &nf1, &bw1, &fl1);
        h->order = ord;
        use_rcm = true;                  // (fl1 / bw1 now describe the caller's ordering)
        n_free_full = nf1; halfbw = bw1;
    }
    std::vector<int> pos(nV);
    for (int p = 0; p < nV; ++p) pos[h->order[p]] = p;

    // lower pairs of every node (static: the solver order does not depend on the subset)
    std::vector<int> lp_ptr(nV + 1, 0), lp_node, lp_eptr(1, 0), lp_elem;
    std::vector<unsigned char> lp_end;
    for (int a = 0; a < nV; ++a) {
        std::vector<std::pair<int, int>> inc;     // (neighbour, incidence slot) of a's incident elements
        for (int p = ne_ptr[a]; p < ne_ptr[a + 1]; ++p) {
            int e = ne_elem[p];
            int c = d->conn[2 * e + (ne_end[p] ? 0 : 1)];
            if (pos[c] < pos[a]) inc.push_back({c, p});
        }
        std::stable_sort(inc.begin(), inc.end(), [](const std::pair<int, int> &x, const std::pair<int, int> &y) {
            return x.first < y.first; });        // per neighbour, elements stay in ascending id
        for (size_t k = 0; k < inc.size(); ++k) {
            if (k == 0 || inc[k].first != inc[k - 1].first) {
                if (k) lp_eptr.push_back((int)lp_elem.size());
                lp_node.push_back(inc[k].first);
            }
            lp_elem.push_back(ne_elem[inc[k].second]);
            lp_end.push_back(ne_end[inc[k].second]);
        }
        if (!inc.empty()) lp_eptr.push_back((int)lp_elem.size());
        lp_ptr[a + 1] = (int)lp_node.size();
    }

    m.n_tile_rows_max = std::max(1, (n_free_full + CM_TILE - 1) / CM_TILE);
    m.band_tiles = (halfbw + CM_TILE - 1) / CM_TILE + 1;
    if (m.band_tiles > m.n_tile_rows_max) m.band_tiles = m.n_tile_rows_max;

    cm_model_info &I = h->info;
    I.n_nodes = nV; I.n_elems = nE; I.n_supports = nS; I.n_loadcases = 0;
    I.mask_words = m.mask_words;
    I.n_free_full = n_free_full;
    I.half_bandwidth = halfbw;
    I.tile = CM_TILE;
    I.n_tile_rows_max = m.n_tile_rows_max;
    I.band_tiles = m.band_tiles;
    I.profile_flops_full = use_rcm ? fl1 : fl0;

    int rc;
    double *props, *cr; int *conn, *d_ne_ptr, *d_ne_elem, *d_sup_node, *d_node_sup, *d_order, *d_pos;
    int *d_lp_ptr, *d_lp_node, *d_lp_eptr, *d_lp_elem;
    unsigned char *d_ne_end, *d_sup_cond, *d_lp_end;
#define UP(expr) do { rc = (expr); if (rc != CM_OK) { cm_destroy(h); return rc; } } while (0)
    UP(dev_upload(h, d->conn, (size_t)2 * nE, &conn));
    UP(dev_upload(h, d->props, (size_t)7 * nE, &props));
    UP(dev_upload(h, d->cr, (size_t)6 * nE, &cr));
    UP(dev_upload(h, ne_ptr.data(), ne_ptr.size(), &d_ne_ptr));
    UP(dev_upload(h, ne_elem.data(), ne_elem.size(), &d_ne_elem));
    UP(dev_upload(h, ne_end.data(), ne_end.size(), &d_ne_end));
    UP(dev_upload(h, d->sup_node, (size_t)nS, &d_sup_node));
    UP(dev_upload(h, d->sup_cond, (size_t)6 * nS, &d_sup_cond));
    UP(dev_upload(h, node_sup.data(), (size_t)nV, &d_node_sup));
    UP(dev_upload(h, h->order.data(), (size_t)nV, &d_order));
    UP(dev_upload(h, pos.data(), (size_t)nV, &d_pos));
    UP(dev_upload(h, lp_ptr.data(), lp_ptr.size(), &d_lp_ptr));
    UP(dev_upload(h, lp_node.data(), lp_node.size(), &d_lp_node));
    UP(dev_upload(h, lp_eptr.data(), lp_eptr.size(), &d_lp_eptr));
    UP(dev_upload(h, lp_elem.data(), lp_elem.size(), &d_lp_elem));
    UP(dev_upload(h, lp_end.data(), lp_end.size(), &d_lp_end));
    {   // K2's index tables as one blob (ints, then bytes, every table 16-byte aligned)
        std::vector<char> blob;
        auto put = [&](int which, const void *src, size_t bytes) {
            blob.resize((blob.size() + 15) & ~(size_t)15);
            m.k2_off[which] = (int)blob.size();
            const char *c = static_cast<const char *>(src);
            blob.insert(blob.end(), c, c + bytes);
        };
        put(CM_K2T_ORDER, h->order.data(), sizeof(int) * nV);
        put(CM_K2T_NE_PTR, ne_ptr.data(), sizeof(int) * ne_ptr.size());
        put(CM_K2T_NE_ELEM, ne_elem.data(), sizeof(int) * ne_elem.size());
        put(CM_K2T_LP_PTR, lp_ptr.data(), sizeof(int) * lp_ptr.size());
        put(CM_K2T_LP_NODE, lp_node.data(), sizeof(int) * lp_node.size());
        put(CM_K2T_LP_EPTR, lp_eptr.data(), sizeof(int) * lp_eptr.size());
        put(CM_K2T_LP_ELEM, lp_elem.data(), sizeof(int) * lp_elem.size());
        put(CM_K2T_NODE_SUP, node_sup.data(), sizeof(int) * nV);
        put(CM_K2T_CONN, d->conn, sizeof(int) * 2 * (size_t)nE);
        put(CM_K2T_NE_END, ne_end.data(), ne_end.size());
        put(CM_K2T_LP_END, lp_end.data(), lp_end.size());
        put(CM_K2T_SUP_COND, d->sup_cond, (size_t)6 * nS);
        blob.resize((blob.size() + 15) & ~(size_t)15);
        char *d_blob;
        UP(dev_upload(h, blob.data(), blob.size(), &d_blob));
        m.k2_blob = d_blob;
        m.k2_blob_bytes = (int)blob.size();
    }
#undef UP
    {   // geometry 0: the model's own node coordinates
        cudaError_t ce = cudaMalloc((void **)&h->d_xyz, (size_t)3 * nV * sizeof(double));
        if (ce == cudaSuccess) ce = cudaMalloc((void **)&h->d_Kg, (size_t)144 * nE * sizeof(double));
        if (ce == cudaSuccess) ce = cudaMalloc((void **)&h->d_Kt, (size_t)144 * nE * sizeof(double));
        if (ce == cudaSuccess) ce = cudaMalloc((void **)&h->d_R3, (size_t)9 * nE * sizeof(double));
        if (ce == cudaSuccess) ce = cudaMalloc((void **)&h->d_elen, (size_t)nE * sizeof(double));
        if (ce == cudaSuccess) ce = cudaMemcpy(h->d_xyz, d->xyz, (size_t)3 * nV * sizeof(double), cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) { std::string msg = std::string("cm_create: ") + cudaGetErrorString(ce); cm_destroy(h); return fail(CM_ERR_CUDA, msg); }
    }
    m.xyz = h->d_xyz; m.Kg = h->d_Kg; m.Kt = h->d_Kt; m.R3 = h->d_R3; m.elen = h->d_elen;
    m.conn = conn; m.props = props; m.cr = cr;
    m.ne_ptr = d_ne_ptr; m.ne_elem = d_ne_elem; m.ne_end = d_ne_end;
    m.sup_node = d_sup_node; m.sup_cond = d_sup_cond; m.node_sup = d_node_sup;
    m.order = d_order; m.pos = d_pos;
    m.lp_ptr = d_lp_ptr; m.lp_node = d_lp_node; m.lp_eptr = d_lp_eptr; m.lp_elem = d_lp_elem; m.lp_end = d_lp_end;
    m.q = nullptr; m.pnodal = nullptr;

    cudaError_t ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) for (int i = 0; i < 4 && ce == cudaSuccess; ++i) ce = cudaEventCreate(&h->ev[i]);
    // (lanes of descending stream priority - so that the waves of a call finish staggered - were measured: slower,
    // 55.3 against 49.0 ms per 12 500 full-result subsets; profiles/r02_experiments.txt)
    for (int i = 0; i < 4 && ce == cudaSuccess; ++i) ce = cudaStreamCreateWithFlags(&h->lane_stream[i], cudaStreamNonBlocking);
    for (int i = 0; i < 5 && ce == cudaSuccess; ++i) ce = cudaEventCreateWithFlags(&h->lane_ev[i], cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 3 && ce == cudaSuccess; ++i) ce = cudaEventCreateWithFlags(&h->aux_ev[i], cudaEventDisableTiming);
    for (int i = 0; i < CM_STAGE_SLOTS && ce == cudaSuccess; ++i) {
        ce = cudaEventCreateWithFlags(&h->slot_done[i], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->slot_free[i], cudaEventDisableTiming);
    }
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->copy_ev, cudaEventDisableTiming);
    for (int i = 0; i < 4 && ce == cudaSuccess; ++i) {
        ce = cudaStreamCreateWithFlags(&h->copy_lane[i], cudaStreamNonBlocking);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->copy_lane_ev[i], cudaEventDisableTiming);
    }
    if (const char *e = getenv("CM_LANES")) { int n = atoi(e); if (n >= 1 && n <= 4) h->n_lanes = n; }
    if (const char *e = getenv("CM_NO_OVERLAP")) h->overlap = (e[0] == '1') ? 0 : 1;
    if (ce == cudaSuccess) ce = cudaEventRecord(h->ev[0], h->stream);
    if (ce == cudaSuccess) ce = cm_launch_element_tables(m, h->stream);
    h->launches++;
    if (ce == cudaSuccess) ce = cudaEventRecord(h->ev[1], h->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
    if (ce == cudaSuccess) { float ms = 0; if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) h->k1_ms = ms; }
    if (ce != cudaSuccess) {
        std::string msg = std::string("cm_create: ") + cudaGetErrorString(ce);
        cm_destroy(h);
        return fail(CM_ERR_CUDA, msg);
    }
    *out = h;
    return CM_OK;
}

int cm_destroy(cm_handle *h) {
    if (!h) return CM_OK;
    cudaSetDevice(h->device);
    for (void *p : h->allocs) cudaFree(p);
    if (h->d_q) cudaFree(h->d_q);
    if (h->d_pnodal) cudaFree(h->d_pnodal);
    cudaFree(h->d_xyz); cudaFree(h->d_Kg); cudaFree(h->d_Kt); cudaFree(h->d_R3); cudaFree(h->d_elen); cudaFree(h->d_geom); cudaFree(h->d_rowoff);
    if (h->stage) cudaFree(h->stage);
    if (h->wsbuf) cudaFree(h->wsbuf);
    for (int i = 0; i < 4; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->stream) cudaStreamDestroy(h->stream);
    for (int i = 0; i < 4; ++i) if (h->lane_stream[i]) cudaStreamDestroy(h->lane_stream[i]);
    for (int i = 0; i < 5; ++i) if (h->lane_ev[i]) cudaEventDestroy(h->lane_ev[i]);
    for (int i = 0; i < CM_STAGE_SLOTS; ++i) {
        if (h->slot_done[i]) cudaEventDestroy(h->slot_done[i]);
        if (h->slot_free[i]) cudaEventDestroy(h->slot_free[i]);
    }
    if (h->copy_ev) cudaEventDestroy(h->copy_ev);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (int i = 0; i < 4; ++i) {
        if (h->copy_lane_ev[i]) cudaEventDestroy(h->copy_lane_ev[i]);
        if (h->copy_lane[i]) cudaStreamDestroy(h->copy_lane[i]);
    }
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    for (int i = 0; i < 3; ++i) if (h->aux_ev[i]) cudaEventDestroy(h->aux_ev[i]);
    delete h;
    return CM_OK;
}

int cm_get_info(const cm_handle *h, cm_model_info *info) {
    if (!h || !info) return fail(CM_ERR_INVALID, "cm_get_info: null argument");
    *info = h->info;
    return CM_OK;
}

int cm_get_element_tables(cm_handle *h, double *Kg, double *R3, int32_t *id_map) {
    if (!h) return fail(CM_ERR_INVALID, "cm_get_element_tables: null handle");
    CU(cudaSetDevice(h->device));
    const int nE = h->m.nE;
    if (Kg) CU(cudaMemcpy(Kg, h->m.Kg, (size_t)144 * nE * sizeof(double), cudaMemcpyDeviceToHost));
    if (R3) CU(cudaMemcpy(R3, h->m.R3, (size_t)9 * nE * sizeof(double), cudaMemcpyDeviceToHost));
    if (id_map) {
        std::vector<int> conn(2 * (size_t)nE);
        CU(cudaMemcpy(conn.data(), h->m.conn, conn.size() * sizeof(int), cudaMemcpyDeviceToHost));
        for (int e = 0; e < nE; ++e)                          // numpy_stiffness.py:567-570
            for (int k = 0; k < 2; ++k)
                for (int i = 0; i < 6; ++i) id_map[12 * (size_t)e + 6 * k + i] = 6 * conn[2 * e + k] + i;
    }
    return CM_OK;
}

int cm_get_node_order(const cm_handle *h, int32_t *order) {
    if (!h || !order) return fail(CM_ERR_INVALID, "cm_get_node_order: null argument");
    std::copy(h->order.begin(), h->order.end(), order);
    return CM_OK;
}

// Per-element property SoA from per-tag tables (host side of cm_create's input): what the reference's
// engine constructor does element by element in Python (numpy_stiffness.py:579-614).
int cm_pack_elements(int32_t n_elems, const int32_t *sec_of, const int32_t *mat_of, const int32_t *joint_of,
                     const uint8_t *bending_stiff, int32_t n_sec, const double *sections, int32_t n_mat,
                     const double *materials, int32_t n_joint, const double *joints, double *props, double *cr) {
    if (n_elems < 0 || !sec_of || !mat_of || !sections || !materials || !props || !cr)
        return fail(CM_ERR_INVALID, "cm_pack_elements: null argument");
    if (n_joint > 0 && !joints) return fail(CM_ERR_INVALID, "cm_pack_elements: joints table missing");
    const double inf = std::numeric_limits<double>::infinity();
    for (int32_t e = 0; e < n_elems; ++e) {
        const int32_t s = sec_of[e], mt = mat_of[e], j = joint_of ? joint_of[e] : -1;
        if (s < 0 || s >= n_sec || mt < 0 || mt >= n_mat || j >= n_joint)
            return fail(CM_ERR_INVALID, "cm_pack_elements: element " + std::to_string(e) + " refers to a table entry that does not exist");
        const double *cs = sections + 4 * (size_t)s, *ma = materials + 3 * (size_t)mt;
        if (!(std::fabs(ma[0]) > CM_EPS) || !(std::fabs(cs[3]) > CM_EPS))            // numpy_stiffness_assembly.py:146
            return fail(CM_ERR_INVALID, "E#" + std::to_string(e) + " - E and Iz must be non-zero");
        double *p = props + 7 * (size_t)e, *c = cr + 6 * (size_t)e;
        p[0] = cs[0]; p[1] = cs[1]; p[2] = cs[2]; p[3] = cs[3]; p[4] = ma[0]; p[5] = ma[1]; p[6] = ma[2];
        for (int i = 0; i < 6; ++i) c[i] = j >= 0 ? joints[6 * (size_t)j + i] : inf;     // None -> rigid (:592-597)
        if (bending_stiff && !bending_stiff[e])
            for (int i = 0; i < 6; ++i) c[i] = 0.0;                                      // truss member (:598-601)
    }
    return CM_OK;
}

int cm_set_geometries(cm_handle *h, int32_t n_geom, const double *xyz) {
    if (!h || !xyz) return fail(CM_ERR_INVALID, "cm_set_geometries: null argument");
    if (n_geom < 1) return fail(CM_ERR_INVALID, "cm_set_geometries: at least one geometry");
    cm_model_dev &m = h->m;
    const size_t nV = m.nV, nE = m.nE, G = (size_t)n_geom;
    for (size_t g = 0; g < G; ++g)                  // numpy_stiffness_assembly.py:288-289, per variant
        for (size_t e = 0; e < nE; ++e) {
            const double *pu = xyz + (g * nV + h->conn[2 * e]) * 3, *pv = xyz + (g * nV + h->conn[2 * e + 1]) * 3;
            const double dx = pu[0] - pv[0], dy = pu[1] - pv[1], dz = pu[2] - pv[2];
            if (!(std::sqrt(dx * dx + dy * dy + dz * dz) >= CM_EPS))
                return fail(CM_ERR_INVALID, "geometry " + std::to_string(g) + ": E#" + std::to_string(e) + " - Bar length too small!");
        }
    CU(cudaSetDevice(h->device));
    // The tables are reused when they are large enough (an optimisation loop calls this once per iterate with the same
    // number of variants: multi-GB allocations cost more than the kernel); otherwise new tables are allocated beside
    // the old ones and swapped in on success.  A failure leaves the previous geometries in place unless the buffers
    // were being reused - then the handle has no valid geometry until the next successful call.
    const bool reuse = (size_t)n_geom <= h->geom_tables_cap;
    double *nx = h->d_xyz, *nK = h->d_Kg, *nT = h->d_Kt, *nR = h->d_R3, *nL = h->d_elen;
    cudaError_t ce = cudaSuccess;
    if (!reuse) {
        nx = nK = nT = nR = nL = nullptr;
        ce = cudaMalloc((void **)&nx, G * 3 * nV * sizeof(double));
        if (ce == cudaSuccess) ce = cudaMalloc((void **)&nK, G * 144 * nE * sizeof(double));
        if (ce == cudaSuccess) ce = cudaMalloc((void **)&nT, G * 144 * nE * sizeof(double));
        if (ce == cudaSuccess) ce = cudaMalloc((void **)&nR, G * 9 * nE * sizeof(double));
        if (ce == cudaSuccess) ce = cudaMalloc((void **)&nL, G * nE * sizeof(double));
    }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(nx, xyz, G * 3 * nV * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (ce == cudaSuccess) {
        cm_model_dev t = m;
        t.n_geom = n_geom; t.xyz = nx; t.Kg = nK; t.Kt = nT; t.R3 = nR; t.elen = nL;
        ce = cudaEventRecord(h->ev[0], h->stream);
        if (ce == cudaSuccess) ce = cm_launch_element_tables(t, h->stream);
        h->launches++;
        if (ce == cudaSuccess) ce = cudaEventRecord(h->ev[1], h->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
        if (ce == cudaSuccess) { float ms = 0; if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) h->k1_ms = ms; }
    }
    if (ce != cudaSuccess) {
        if (!reuse) { cudaFree(nx); cudaFree(nK); cudaFree(nT); cudaFree(nR); cudaFree(nL); }
        else { m.n_lc = 0; h->info.n_loadcases = 0; }
        return fail(ce == cudaErrorMemoryAllocation ? CM_ERR_NOMEM : CM_ERR_CUDA, std::string("cm_set_geometries: ") + cudaGetErrorString(ce));
    }
    if (!reuse) {
        cudaFree(h->d_xyz); cudaFree(h->d_Kg); cudaFree(h->d_Kt); cudaFree(h->d_R3); cudaFree(h->d_elen);
        h->d_xyz = nx; h->d_Kg = nK; h->d_Kt = nT; h->d_R3 = nR; h->d_elen = nL;
        h->geom_tables_cap = (size_t)n_geom;
    }
    m.n_geom = n_geom; m.xyz = nx; m.Kg = nK; m.Kt = nT; m.R3 = nR; m.elen = nL;
    // the lumped element loads depend on the geometry: the load cases have to be set again
    if (h->d_q) { cudaFree(h->d_q); h->d_q = nullptr; }
    m.q = nullptr;
    m.n_lc = 0;
    h->info.n_loadcases = 0;
    return CM_OK;
}

int cm_get_geometry_tables(cm_handle *h, int32_t g, double *Kg, double *R3) {
    if (!h) return fail(CM_ERR_INVALID, "cm_get_geometry_tables: null handle");
    if (g < 0 || g >= h->m.n_geom) return fail(CM_ERR_INVALID, "cm_get_geometry_tables: no such geometry");
    CU(cudaSetDevice(h->device));
    const size_t nE = h->m.nE;
    if (Kg) CU(cudaMemcpy(Kg, h->m.Kg + (size_t)g * 144 * nE, 144 * nE * sizeof(double), cudaMemcpyDeviceToHost));
    if (R3) CU(cudaMemcpy(R3, h->m.R3 + (size_t)g * 9 * nE, 9 * nE * sizeof(double), cudaMemcpyDeviceToHost));
    return CM_OK;
}

int32_t cm_geometry_count(const cm_handle *h) { return h ? h->m.n_geom : 0; }

int cm_last_element_kernel_ms(const cm_handle *h, double *ms) {
    if (!h || !ms) return fail(CM_ERR_INVALID, "cm_last_element_kernel_ms: null argument");
    *ms = h->k1_ms;
    return CM_OK;
}

static void make_layout(cm_handle *h) {
    const cm_model_dev &m = h->m;
    cm_ws_layout &L = h->L;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const int n_lc = std::max(1, m.n_lc);
    L.n_pad = m.n_tile_rows_max * CM_TILE;
    size_t o = 0;
    L.off_tiles = o;  o = al(o + (size_t)m.n_tile_rows_max * m.band_tiles * CM_TILE_ELEMS * sizeof(double));
    L.off_sidx = o;   o = al(o + (size_t)6 * m.nV * sizeof(int));
    L.off_dinv = o;   o = al(o + (size_t)L.n_pad * sizeof(double));
    L.off_dpiv = o;   o = al(o + (size_t)(L.n_pad + 8) * sizeof(double));
    L.off_x = o;      o = al(o + (size_t)n_lc * L.n_pad * sizeof(double));
    L.off_p = o;      o = al(o + (size_t)n_lc * 6 * m.nV * sizeof(double));
    L.off_kfirst = o; o = al(o + (size_t)m.n_tile_rows_max * sizeof(int));
    L.off_minv = o;   o = al(o + (size_t)m.n_tile_rows_max * CM_TILE_ELEMS * sizeof(double));
    L.off_kcol = o;   o = al(o + (size_t)m.n_tile_rows_max * sizeof(int));
    L.off_aflag = o;  o = al(o + (size_t)m.n_tile_rows_max * m.band_tiles);
    L.off_info = o;   o = al(o + CM_INFO_INTS * sizeof(int));
    L.per_subset = o;
}

int cm_set_loadcases(cm_handle *h, int32_t n_lc, const double *nodal, const double *w_udl,
                     const uint8_t *has_udl, const uint8_t *grav_on, const double *grav_dir) {
    if (!h) return fail(CM_ERR_INVALID, "cm_set_loadcases: null handle");
    if (n_lc < 1 || n_lc > CM_MAX_LC) return fail(CM_ERR_INVALID, "cm_set_loadcases: 1..8 load cases supported");
    if (!nodal || !grav_on || !grav_dir) return fail(CM_ERR_INVALID, "cm_set_loadcases: null argument");
    if (has_udl && !w_udl) return fail(CM_ERR_INVALID, "cm_set_loadcases: has_udl without w_udl");
    CU(cudaSetDevice(h->device));
    cm_model_dev &m = h->m;
    const size_t nE = m.nE, nV = m.nV;
    // New tables are built beside the old ones and swapped in only when every step succeeded: a
    // failure leaves the handle exactly as it was (old load cases, old layout).
    double *nq = nullptr, *npn = nullptr;
    double *d_w = nullptr, *d_g = nullptr;
    unsigned char *d_has = nullptr, *d_on = nullptr;
    cudaError_t ce = cudaMalloc((void **)&nq, (size_t)m.n_geom * n_lc * nE * 12 * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMalloc((void **)&npn, n_lc * 6 * nV * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMemcpy(npn, nodal, n_lc * 6 * nV * sizeof(double), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess && has_udl) {
        ce = cudaMalloc((void **)&d_w, n_lc * nE * 3 * sizeof(double));
        if (ce == cudaSuccess) ce = cudaMalloc((void **)&d_has, n_lc * nE);
        if (ce == cudaSuccess) ce = cudaMemcpy(d_w, w_udl, n_lc * nE * 3 * sizeof(double), cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) ce = cudaMemcpy(d_has, has_udl, n_lc * nE, cudaMemcpyHostToDevice);
    }
    if (ce == cudaSuccess) ce = cudaMalloc((void **)&d_g, n_lc * 3 * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMalloc((void **)&d_on, n_lc);
    if (ce == cudaSuccess) ce = cudaMemcpy(d_g, grav_dir, n_lc * 3 * sizeof(double), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(d_on, grav_on, n_lc, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) {
        ce = cm_launch_element_loads(m, n_lc, d_w, d_has, d_on, d_g, nq, h->stream);
        h->launches++;
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
    cudaFree(d_w); cudaFree(d_has); cudaFree(d_g); cudaFree(d_on);
    if (ce != cudaSuccess) {
        cudaFree(nq); cudaFree(npn);
        return fail(CM_ERR_CUDA, std::string("cm_set_loadcases: ") + cudaGetErrorString(ce));
    }
    if (h->d_q) cudaFree(h->d_q);
    if (h->d_pnodal) cudaFree(h->d_pnodal);
    h->d_q = nq;
    h->d_pnodal = npn;
    m.n_lc = n_lc;
    m.q = h->d_q;
    m.pnodal = h->d_pnodal;
    h->info.n_loadcases = n_lc;
    make_layout(h);
    return CM_OK;
}

int cm_get_element_loads(cm_handle *h, int32_t lc, double *q) {
    if (!h || !q) return fail(CM_ERR_INVALID, "cm_get_element_loads: null argument");
    if (lc < 0 || lc >= h->m.n_lc) return fail(CM_ERR_STATE, "cm_get_element_loads: load case not set");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpy(q, h->d_q + (size_t)lc * h->m.nE * 12, (size_t)h->m.nE * 12 * sizeof(double), cudaMemcpyDeviceToHost));
    return CM_OK;
}

int cm_set_tol(cm_handle *h, double trans_tol, double rot_tol) {
    if (!h) return fail(CM_ERR_INVALID, "cm_set_tol: null handle");
    h->m.trans_tol = trans_tol;
    h->rot_tol = rot_tol;
    return CM_OK;
}

size_t cm_workspace_bytes(const cm_handle *h, int64_t n) {
    if (!h || h->m.n_lc == 0 || n <= 0) return 0;
    return (size_t)n * h->L.per_subset;
}

int cm_set_variant(cm_handle *h, int32_t variant) {     // 0 scalar validation, 2/3 tensor-core kernel (4/8 warps per CTA), 4/5 instrumented
    if (!h) return fail(CM_ERR_INVALID, "cm_set_variant: null handle");
    if (variant < 0 || variant > 13) return fail(CM_ERR_INVALID, "cm_set_variant: variant must be 0..13 (7/8/9: 2-/3-/1-warp CTAs; 10/11/12: operand ring; 13: tensor-memory look-ahead)");
    h->variant = variant;
    return CM_OK;
}

int cm_debug_k3_phase_cycles(cm_handle *h, uint64_t *out16) {   // variant 4 (instrumented) counters
    if (!h || !out16) return fail(CM_ERR_INVALID, "cm_debug_k3_phase_cycles: null argument");
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    CU(cm_factor_flow_read_prof(reinterpret_cast<unsigned long long *>(out16)));
    return CM_OK;
}

int cm_debug_k2_phase_cycles(cm_handle *h, uint64_t *out8) {   // 16 slots    // phase counters of a -DK2_PROF=1 build of the assembly kernel (zeros otherwise)
    if (!h || !out8) return fail(CM_ERR_INVALID, "cm_debug_k2_phase_cycles: null argument");
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    CU(cm_k2_read_prof(reinterpret_cast<unsigned long long *>(out8)));
    return CM_OK;
}

}  // extern "C"

// ---- wave plan.  The batch is cut into equal waves (a short last wave would leave most SMs idle).
// Outside timing mode, and when the batch needs more than one wave, the workspace is split into
// n_lanes parts and consecutive waves rotate over as many internal streams forked from / joined to
// the caller's stream: a factorisation launch drains over roughly one CTA lifetime, and the other
// lanes' kernels fill the SMs it frees.  Every wave is independent (own subsets, own outputs, own
// share of the workspace).
struct WavePlan { bool lanes; int NL; long long cap, n_waves, wsize; std::vector<int> sizes; };      // sizes: empty = equal waves of wsize
// max_wave > 0: upper bound for the subsets of one wave (the host entry point with large outputs wants waves whose
// device->host copy is short, see cm_solve_many_host_g)
static int plan_waves(const cm_handle *h, int64_t B, size_t ws_bytes, WavePlan *p, long long max_wave = 0, size_t out_per_subset = 0) {
    const size_t per = h->L.per_subset;
    long long wave = (long long)(ws_bytes / per);
    if (wave < 1) return fail(CM_ERR_NOMEM, "cm_solve_many: workspace smaller than one subset (" + std::to_string(per) + " bytes)");
    if (wave > B) wave = B;
    const int sm_fill = getenv("CM_LANE_MIN") ? atoi(getenv("CM_LANE_MIN")) : 8 * h->m.n_sm;     // two full sets of resident factorisation CTAs
    p->NL = h->n_lanes;
    p->lanes = !h->timing && h->overlap && p->NL > 1 && h->lane_stream[0] && B > wave / p->NL && wave / p->NL >= sm_fill;
    p->cap = p->lanes ? wave / p->NL : wave;
    if (max_wave > 0 && p->cap > std::max<long long>(max_wave, sm_fill)) p->cap = std::max<long long>(max_wave, sm_fill);
    p->n_waves = (B + p->cap - 1) / p->cap;
    if (p->lanes && getenv("CM_WAVE_ROUND") && atoi(getenv("CM_WAVE_ROUND")) && p->n_waves % p->NL) p->n_waves += p->NL - p->n_waves % p->NL;   // experiment: same number of waves per lane
    p->wsize = (B + p->n_waves - 1) / p->n_waves;
    p->sizes.clear();
    // Large outputs (max_wave > 0): the results of a wave exist when its last kernel has finished, so whatever the last
    // waves of the lanes produce is copied to the host AFTER the kernels have ended - an unhidden tail of about three
    // waves' results with equal waves (5 ms of a 49 ms call on the 1k-element frame).  The waves therefore shrink
    // geometrically towards the end of the batch (by 0.8 per wave down to 4/3 CTAs per SM; 70 % / two per SM measured
    // 46.6, 80 % / 1.35 per SM 46.3 ms per 12 500 full-result subsets, equal waves 48.6): what is still to be copied when
    // the last kernel ends is small, and the waves in front of them stay large.
    static const int taper_pct = getenv("CM_TAPER") ? atoi(getenv("CM_TAPER")) : 80;
    static const int taper_min = getenv("CM_TAPER_MIN") ? atoi(getenv("CM_TAPER_MIN")) : 0;
    static const int stagger = getenv("CM_STAGGER") ? atoi(getenv("CM_STAGGER")) : 0;     // first waves of the lanes 1/NL, 2/NL, ... of a wave
    if (max_wave > 0 && p->lanes && p->n_waves > 2 * p->NL && ((taper_pct > 0 && taper_pct < 100) || stagger)) {
        // smallest wave: about half a millisecond of device->host copy (26 MB), not below 4/3 CTAs per SM
        const long long wmin = taper_min > 0 ? taper_min
                                             : std::max<long long>(4 * h->m.n_sm / 3, out_per_subset ? (26ll << 20) / (long long)out_per_subset : 0);
        std::vector<int> head, tail;             // tail: from the end of the batch backwards
        long long t_sum = 0;
        if (stagger)
            for (int i = 1; i < p->NL; ++i) { head.push_back((int)std::max<long long>(wmin, p->wsize * i / p->NL)); t_sum += head.back(); }
        if (taper_pct > 0 && taper_pct < 100)
            for (double v = (double)wmin; v < (double)p->wsize && t_sum + (long long)v < B / 2; v = v * 100.0 / taper_pct) {
                tail.push_back((int)v);
                t_sum += (long long)v;
            }
        const long long rest = B - t_sum;
        const long long n_full = (rest + p->wsize - 1) / p->wsize;
        const long long full = (rest + n_full - 1) / n_full;
        p->sizes = head;
        for (long long r = rest; r > 0; r -= full) p->sizes.push_back((int)std::min<long long>(full, r));
        for (size_t i = tail.size(); i-- > 0;) p->sizes.push_back(tail[i]);
        p->n_waves = (long long)p->sizes.size();
    }
    return CM_OK;
}

static cm_out_dev out_dev_of(const cm_solve_out *o) {
    cm_out_dev od;
    od.flag = reinterpret_cast<unsigned char *>(o->flag);
    od.status = reinterpret_cast<unsigned char *>(o->status);
    od.n_free = reinterpret_cast<int *>(o->n_free);
    od.n_fixed = reinterpret_cast<int *>(o->n_fixed);
    od.dof_map = reinterpret_cast<int *>(o->dof_map);
    od.node_exists = reinterpret_cast<unsigned char *>(o->node_exists);
    od.U = reinterpret_cast<double *>(o->U);
    od.Rf = reinterpret_cast<double *>(o->Rf);
    od.eR = reinterpret_cast<double *>(o->eR);
    od.compliance = reinterpret_cast<double *>(o->compliance);
    od.max_trans = reinterpret_cast<double *>(o->max_trans);
    od.max_rot = reinterpret_cast<double *>(o->max_rot);
    od.min_pivot = reinterpret_cast<double *>(o->min_pivot);
    od.profile_flops = reinterpret_cast<double *>(o->profile_flops);
    od.sigma_max = reinterpret_cast<double *>(o->sigma_max);
    od.max_abs_sigma = reinterpret_cast<double *>(o->max_abs_sigma);
    od.profile_entries = reinterpret_cast<double *>(o->profile_entries);
    od.envelope_tiles = reinterpret_cast<double *>(o->envelope_tiles);
    od.Rf_sup = reinterpret_cast<double *>(o->Rf_sup);
    od.eR_rows = reinterpret_cast<double *>(o->eR_rows);
    od.eR_row_offsets = reinterpret_cast<const long long *>(o->eR_row_offsets);
    od.row_bias = 0;             // (device entry points: off[0] == 0)
    return od;
}

// The waves of one call.  `before(w, first, count, lane_stream, &od)` runs before a wave's kernels are
// enqueued and may redirect the wave's outputs; `after(...)` runs right after its last kernel.
template <typename Before, typename After>
static int run_waves(cm_handle *h, const uint32_t *masks, const int *geom, const int *trows, int group, int64_t B,
                     const cm_out_dev &od0, char *ws, const WavePlan &P, cudaStream_t st, Before before, After after) {
    const size_t per = h->L.per_subset;
    double acc_ms[3] = {0, 0, 0};
    if (P.lanes) {
        CU(cudaEventRecord(h->lane_ev[0], st));
        for (int i = 0; i < P.NL; ++i) CU(cudaStreamWaitEvent(h->lane_stream[i], h->lane_ev[0], 0));
    }
    long long w = 0;
    for (long long first = 0, step = 0; first < B; first += step, ++w) {
        step = P.sizes.empty() ? P.wsize : P.sizes[(size_t)w];
        const int count = (int)std::min<long long>(step, B - first);
        cudaStream_t ls = P.lanes ? h->lane_stream[w % P.NL] : st;
        char *lws = P.lanes ? ws + (size_t)(w % P.NL) * (size_t)P.cap * per : ws;
        cm_out_dev od = od0;
        { int rc = before(w, first, count, ls, od); if (rc != CM_OK) return rc; }
        const int vcfg = h->variant >= 10 ? h->variant :
                         h->variant == 1 ? 0 : (h->variant >= 7 ? h->variant - 3 : (h->variant == 6 ? 3 : ((h->variant & 1) ? 1 : 2)));   // 7/8/9: 2/3/1 warps; 10-12: operand ring
        const int prof = (h->variant == 4 || h->variant == 5) ? 1 : 0;
        const int ktrm = h->variant != 0 ? 1 : 0;
        if (h->timing) CU(cudaEventRecord(h->ev[0], ls));
        if (trows && group > 1 && h->variant != 0) {
            // Prefix-sharing groups.  The trunks (first subset of every group) are assembled and factored first - few
            // CTAs that fill a fraction of the GPU - while a second stream assembles all the other subsets; then the
            // wave is ranked and the others are factored, reading the trunks' rows.
            const int n_trunks = (count + group - 1) / group;
            const bool two = !h->timing && h->aux_stream;
            cudaStream_t bs = two ? h->aux_stream : ls;
            if (two) { CU(cudaEventRecord(h->aux_ev[0], ls)); CU(cudaStreamWaitEvent(bs, h->aux_ev[0], 0)); }
            CU(cm_launch_dofmap_assemble(h->m, h->L, lws, masks, geom, trows, group, 1, first, count, od, ktrm, ls));
            if (two) CU(cudaEventRecord(h->aux_ev[1], ls));
            if (!h->timing) CU(cm_launch_factor_flow(h->m, h->L, lws, 0, n_trunks, group, vcfg, prof, ls));
            CU(cm_launch_dofmap_assemble(h->m, h->L, lws, masks, geom, trows, group, 2, first, count, od, ktrm, bs));
            if (two) CU(cudaStreamWaitEvent(bs, h->aux_ev[1], 0));              // the ranking reads the trunks' cost keys as well
            CU(cm_launch_schedule(h->L, lws, count, bs));
            if (two) { CU(cudaEventRecord(h->aux_ev[2], bs)); CU(cudaStreamWaitEvent(ls, h->aux_ev[2], 0)); }
            if (h->timing) {
                CU(cudaEventRecord(h->ev[1], ls));
                CU(cm_launch_factor_flow(h->m, h->L, lws, 0, n_trunks, group, vcfg, prof, ls));
            }
            if (count > n_trunks) CU(cm_launch_factor_flow(h->m, h->L, lws, n_trunks, count - n_trunks, 0, vcfg, prof, ls));
            h->launches += 3;
        } else {
            CU(cm_launch_dofmap_assemble(h->m, h->L, lws, masks, geom, nullptr, 0, 0, first, count, od, ktrm, ls));
            if (h->variant != 0) { CU(cm_launch_schedule(h->L, lws, count, ls)); h->launches += 1; }
            if (h->timing) CU(cudaEventRecord(h->ev[1], ls));
            if (h->variant == 0)
                CU(cm_launch_factor(h->m, h->L, lws, count, 0, ls));
            else
                CU(cm_launch_factor_flow(h->m, h->L, lws, 0, count, 0, vcfg, prof, ls));
        }
        if (h->timing) CU(cudaEventRecord(h->ev[2], ls));
        CU(cm_launch_solve_recover(h->m, h->L, lws, masks, geom, first, count, od, ls));
        h->launches += 3;
        if (h->timing) {
            CU(cudaEventRecord(h->ev[3], ls));
            CU(cudaEventSynchronize(h->ev[3]));
            for (int i = 0; i < 3; ++i) {
                float ms = 0;
                CU(cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]));
                acc_ms[i] += ms;
            }
        }
        { int rc = after(w, first, count, ls); if (rc != CM_OK) return rc; }
    }
    if (P.lanes)
        for (int i = 0; i < P.NL; ++i) {
            CU(cudaEventRecord(h->lane_ev[1 + i], h->lane_stream[i]));
            CU(cudaStreamWaitEvent(st, h->lane_ev[1 + i], 0));
        }
    if (h->timing) for (int i = 0; i < 3; ++i) h->stage_ms[i] = acc_ms[i];
    h->last_waves = (int)P.n_waves;
    return CM_OK;
}

extern "C" {

int cm_solve_many(cm_handle *h, uint64_t masks_dev, int64_t B, const cm_solve_out *o, uint64_t ws_dev,
                  size_t ws_bytes, uint64_t stream) {
    return cm_solve_many_g(h, masks_dev, 0, B, o, ws_dev, ws_bytes, stream);
}

int cm_solve_many_g(cm_handle *h, uint64_t masks_dev, uint64_t geom_dev, int64_t B, const cm_solve_out *o, uint64_t ws_dev,
                    size_t ws_bytes, uint64_t stream) {
    if (!h || !o) return fail(CM_ERR_INVALID, "cm_solve_many: null argument");
    if (h->m.n_lc == 0) return fail(CM_ERR_STATE, "cm_solve_many: call cm_set_loadcases first");
    if (B <= 0) return CM_OK;
    if (!masks_dev || !ws_dev) return fail(CM_ERR_INVALID, "cm_solve_many: null device buffer");
    if (o->eR_rows && !o->eR_row_offsets) return fail(CM_ERR_INVALID, "cm_solve_many: eR_rows needs eR_row_offsets");
    WavePlan P;
    { int rc = plan_waves(h, B, ws_bytes, &P); if (rc != CM_OK) return rc; }
    CU(cudaSetDevice(h->device));
    return run_waves(h, reinterpret_cast<const uint32_t *>(masks_dev), reinterpret_cast<const int *>(geom_dev), nullptr, 0, B, out_dev_of(o),
                     reinterpret_cast<char *>(ws_dev), P,
                     reinterpret_cast<cudaStream_t>(stream),
                     [](long long, long long, int, cudaStream_t, cm_out_dev &) { return CM_OK; },
                     [](long long, long long, int, cudaStream_t) { return CM_OK; });
}

}  // extern "C"

// per-subset bytes of every output of cm_solve_out_host, in the order of the struct
namespace {
inline int64_t popcount_words(const uint32_t *w, int n) {
    int64_t cnt = 0;
    int i = 0;
    for (; i + 1 < n; i += 2) cnt += __builtin_popcountll((unsigned long long)w[i] | ((unsigned long long)w[i + 1] << 32));
    if (i < n) cnt += __builtin_popcount(w[i]);
    return cnt;
}
struct OutItem { void *host; size_t stride; size_t off; };       // stride = bytes per subset; off = offset inside a staging slot
constexpr int N_OUT = 19;      // fixed-stride outputs; eR_rows (rows of the existing elements only) is handled beside them
void out_items(const cm_handle *h, const cm_solve_out_host *o, OutItem (&it)[N_OUT]) {
    const size_t nV = h->m.nV, nE = h->m.nE, Lc = h->m.n_lc;
    const OutItem t[N_OUT] = {
        {o->flag, Lc, 0}, {o->status, Lc, 0}, {o->n_free, 4, 0}, {o->n_fixed, 4, 0}, {o->dof_map, 6 * nV * 4, 0},
        {o->node_exists, nV, 0}, {o->U, Lc * 6 * nV * 8, 0}, {o->Rf, Lc * 6 * nV * 8, 0}, {o->eR, Lc * nE * 12 * 8, 0},
        {o->compliance, Lc * 8, 0}, {o->max_trans, Lc * 8, 0}, {o->max_rot, Lc * 8, 0}, {o->min_pivot, 8, 0},
        {o->profile_flops, 8, 0}, {o->sigma_max, Lc * nE * 8, 0}, {o->max_abs_sigma, Lc * 8, 0},
        {o->profile_entries, 8, 0}, {o->envelope_tiles, 8, 0}, {o->Rf_sup, Lc * (size_t)h->m.nS * 6 * 8, 0}};
    for (int i = 0; i < N_OUT; ++i) it[i] = t[i];
}
void out_advance(const cm_handle *h, cm_solve_out_host *o, int64_t b0) {
    OutItem it[N_OUT];
    out_items(h, o, it);
    void **f[N_OUT] = {(void **)&o->flag, (void **)&o->status, (void **)&o->n_free, (void **)&o->n_fixed, (void **)&o->dof_map,
                       (void **)&o->node_exists, (void **)&o->U, (void **)&o->Rf, (void **)&o->eR, (void **)&o->compliance,
                       (void **)&o->max_trans, (void **)&o->max_rot, (void **)&o->min_pivot, (void **)&o->profile_flops,
                       (void **)&o->sigma_max, (void **)&o->max_abs_sigma, (void **)&o->profile_entries,
                       (void **)&o->envelope_tiles, (void **)&o->Rf_sup};
    for (int i = 0; i < N_OUT; ++i)
        if (it[i].host) *f[i] = static_cast<char *>(it[i].host) + (size_t)b0 * it[i].stride;
    if (o->eR_row_offsets) {          // subset b0's rows start (off[b0] - off[0]) rows into eR_rows
        if (o->eR_rows) o->eR_rows += (size_t)(o->eR_row_offsets[b0] - o->eR_row_offsets[0]) * h->m.n_lc * 12;
        o->eR_row_offsets += b0;
    }
}
}  // namespace

extern "C" {

static int solve_host(cm_handle *h, const uint32_t *masks_host, const int32_t *geom_host, const int32_t *trows_host, int group,
                      int64_t B, const cm_solve_out_host *o, int64_t chunk = 0, const volatile int32_t *ready = nullptr);

// Host buffers.  Masks go up in one copy.  Outputs are staged per WAVE in a ring of device slots: right
// after a wave's last kernel its slot is drained to the caller's arrays by a copy stream (one
// cudaMemcpyAsync per requested output, written at the subset's final position), while the lanes
// compute the following waves; a slot is reused only after its copies have finished.  The device
// staging is therefore bounded by CM_STAGE_SLOTS waves, not by the batch, and the device->host
// transfer of full results (U, Rf, eR: 127 KB per subset on the 1k-element frame) overlaps the
// kernels instead of following them.
int cm_solve_many_host(cm_handle *h, const uint32_t *masks_host, int64_t B, const cm_solve_out_host *o) {
    return cm_solve_many_host_g(h, masks_host, nullptr, B, o);
}

int cm_solve_many_host_g(cm_handle *h, const uint32_t *masks_host, const int32_t *geom_host, int64_t B,
                         const cm_solve_out_host *o) {
    return solve_host(h, masks_host, geom_host, nullptr, 0, B, o);
}

// Prefix-sharing groups (the construction-sequence path).  Consecutive subsets of one group
// are prefixes of ONE construction sequence: in the handle's static node ordering the matrix of a later prefix differs
// from the first subset of its group (the "trunk") only at and behind the first solver row that an added element
// touches or a new node is inserted in front of.  The trunk is factored completely; a "tail" subset b factors only the
// tile rows >= t_rows[b] and reads the rows above from the trunk's slab (tiles of L, M_J, pivots, forward solutions) -
// in the factorisation AND in the back substitution.  Waves hold whole groups; the trunks are assembled and factored
// first, beside the assembly of the tails on a second stream.
int cm_solve_many_host_seq(cm_handle *h, const uint32_t *masks_host, int64_t B, int32_t group, const int32_t *t_rows_host,
                           const cm_solve_out_host *o) {
    if (!t_rows_host || group < 1) return fail(CM_ERR_INVALID, "cm_solve_many_host_seq: group >= 1 and t_rows required");
    if (h && h->variant == 0) return fail(CM_ERR_STATE, "cm_solve_many_host_seq: not available with the scalar validation kernel");
    for (int64_t b = 0; b < B; ++b)
        if (t_rows_host[b] < 0 || (b % group == 0 && t_rows_host[b] != 0))
            return fail(CM_ERR_INVALID, "cm_solve_many_host_seq: t_rows must be >= 0, and 0 for the first subset of every group");
    return solve_host(h, masks_host, nullptr, t_rows_host, group, B, o);
}

}  // extern "C"

// chunk / ready (cm_solve_many_host_chunks): the masks arrive in consecutive chunks of `chunk` subsets, chunk k is complete
// once ready[k] != 0 (< 0: cancelled).  A wave is enqueued as soon as the chunks it covers are there and takes its own
// masks to the device - the producer (the Python facade packing id lists) runs beside the GPU, wave by wave.
static int solve_host(cm_handle *h, const uint32_t *masks_host, const int32_t *geom_host, const int32_t *trows_host, int group,
                      int64_t B, const cm_solve_out_host *o, int64_t chunk, const volatile int32_t *ready) {
    if (!h || !o || !masks_host) return fail(CM_ERR_INVALID, "cm_solve_many_host: null argument");
    if (ready && (chunk <= 0 || o->eR_rows || trows_host))
        return fail(CM_ERR_INVALID, "cm_solve_many_host_chunks: chunk must be positive; compact element rows and sequences need all masks up front");
    if (h->m.n_lc == 0) return fail(CM_ERR_STATE, "cm_solve_many_host: call cm_set_loadcases first");
    if (B <= 0) return CM_OK;
    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
    double host_mark[4] = {0, 0, 0, 0};
    if (geom_host)
        for (int64_t b = 0; b < B; ++b)
            if (geom_host[b] < 0 || geom_host[b] >= h->m.n_geom)
                return fail(CM_ERR_INVALID, "cm_solve_many_host: geometry index " + std::to_string(geom_host[b]) + " of subset " +
                                            std::to_string(b) + " not in [0, " + std::to_string(h->m.n_geom) + ")");
    const int64_t *roff = o->eR_rows ? o->eR_row_offsets : nullptr;
    if (o->eR_rows) {
        // the caller sized eR_rows from its offsets: they must be the exclusive prefix sum of the masks' popcounts, or the
        // kernel would write rows where the caller did not reserve any
        if (!roff) return fail(CM_ERR_INVALID, "cm_solve_many_host: eR_rows needs eR_row_offsets (i64 [B+1])");
        const int W = h->m.mask_words;
        for (int64_t b = 0; b < B; ++b) {
            const int64_t cnt = popcount_words(masks_host + (size_t)b * W, W);
            if (roff[b + 1] - roff[b] != cnt)
                return fail(CM_ERR_INVALID, "cm_solve_many_host: eR_row_offsets[" + std::to_string(b + 1) + "] - [" + std::to_string(b) +
                                            "] is not the number of elements of subset " + std::to_string(b));
        }
    }
    CU(cudaSetDevice(h->device));
    if (roff && (size_t)B + 1 > h->rowoff_cap) {
        cudaFree(h->d_rowoff);
        h->d_rowoff = nullptr; h->rowoff_cap = 0;
        CU(cudaMalloc((void **)&h->d_rowoff, ((size_t)B + 1) * sizeof(long long)));
        h->rowoff_cap = (size_t)B + 1;
    }
    const int32_t *aux_host = geom_host ? geom_host : trows_host;       // one per-subset int array at a time
    if (aux_host && (size_t)B > h->geom_cap) {
        cudaFree(h->d_geom);
        h->d_geom = nullptr; h->geom_cap = 0;
        CU(cudaMalloc((void **)&h->d_geom, (size_t)B * sizeof(int)));
        h->geom_cap = (size_t)B;
    }
    if (trows_host && 2 * (size_t)h->L.n_pad * sizeof(double) > 100 * 1024)
        return fail(CM_ERR_INVALID, "cm_solve_many_host_seq: model too large for shared rows (solution vectors must fit shared memory)");
    const cm_model_dev &m = h->m;
    const size_t Bn = (size_t)B;
    // device staging: the masks of the batch, then CM_STAGE_SLOTS slots of one wave's outputs each
    OutItem it[N_OUT];
    out_items(h, o, it);
    // A wave's results leave for the host when its last kernel has finished, and the copies of consecutive waves
    // queue on one copy engine: with full results (127 KB per subset on the 1k-element frame) a wave of 2500 subsets
    // is a 6 ms copy, and the last waves of all lanes finish together - their copies are an unhidden tail.  Large
    // outputs therefore run in waves of about 64 MB of results (never below what fills the GPU twice).
    size_t out_per_subset = 0;
    for (int i = 0; i < N_OUT; ++i) if (it[i].host) out_per_subset += it[i].stride;
    if (roff) out_per_subset += (size_t)((roff[B] - roff[0]) / B) * m.n_lc * 96;
    static const long long wave_mb = getenv("CM_HOST_WAVE_MB") ? atoll(getenv("CM_HOST_WAVE_MB")) : 64;
    long long max_wave = out_per_subset > 16384 ? std::max<long long>(1, (wave_mb << 20) / (long long)out_per_subset) : 0;
    if (ready && max_wave == 0) max_wave = 1;      // masks still being produced: small waves (two GPU fills) start early
    // workspace: enough subsets in flight to fill the GPU several times over, bounded by memory
    // (the driver is only asked about free memory when the workspace has to grow)
    const size_t per = h->L.per_subset;
    static const size_t ws_subsets = getenv("CM_WS_SUBSETS") ? (size_t)std::max(1, atoi(getenv("CM_WS_SUBSETS"))) : 8192;      // 4096: 266 k, 8192: 273 k, 12500: 272 k solves/s on config 3 (profiles/r01c_experiments.txt)
    size_t want_subsets = std::min<size_t>(Bn, ws_subsets);
    if (max_wave > 0 && !(trows_host && group > 1)) {
        // bounded waves (large outputs, chunked masks): no more slabs than the lanes' waves use (19 GB instead of 45 on the 1k frame)
        const long long sm_fill = getenv("CM_LANE_MIN") ? atoi(getenv("CM_LANE_MIN")) : 8 * h->m.n_sm;
        want_subsets = std::min<size_t>(want_subsets, (size_t)h->n_lanes * (size_t)std::max<long long>(max_wave, sm_fill));
    }
    if (trows_host && group > 1) want_subsets = (want_subsets + group - 1) / group * group;     // whole groups
    size_t want = want_subsets * per;
    if (want > h->ws_bytes) {
        size_t freeb = 0, totalb = 0;
        CU(cudaMemGetInfo(&freeb, &totalb));
        size_t cap = (freeb + h->ws_bytes) / 2;
        if (want > cap) want = std::max(per, cap / per * per);
    }
    if (want > h->ws_bytes) {
        if (h->wsbuf) cudaFree(h->wsbuf);
        h->wsbuf = nullptr; h->ws_bytes = 0;
        CU(cudaMalloc((void **)&h->wsbuf, want));
        h->ws_bytes = want;
    }
    WavePlan P;
    { int rc = plan_waves(h, B, h->ws_bytes, &P, max_wave, out_per_subset); if (rc != CM_OK) return rc; }
    if (trows_host && group > 1) {
        // waves hold whole groups and run one after the other on the caller-facing stream: the overlap inside a wave
        // (trunk factorisation beside the assembly of the tails) takes the place of the lanes
        const long long cap_groups = (long long)(h->ws_bytes / h->L.per_subset) / group;
        if (cap_groups < 1) return fail(CM_ERR_NOMEM, "cm_solve_many_host_seq: workspace smaller than one group");
        const long long n_groups = (B + group - 1) / group;
        const long long n_w = (n_groups + cap_groups - 1) / cap_groups;
        P.lanes = false;
        P.sizes.clear();
        P.n_waves = n_w;
        P.wsize = (n_groups + n_w - 1) / n_w * group;
        P.cap = P.wsize;
    }
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t mask_bytes = al(Bn * m.mask_words * 4);
    size_t slot_bytes = 0;
    for (int i = 0; i < N_OUT; ++i)
        if (it[i].host) { it[i].off = slot_bytes; slot_bytes = al(slot_bytes + (size_t)P.wsize * it[i].stride); }
    size_t rows_off = 0;                 // the compact element rows of a wave: sized by the wave with the most rows
    if (roff) {
        long long max_rows = 0, first = 0;
        for (long long w = 0; first < B; ++w) {
            const long long cnt = std::min<long long>(P.sizes.empty() ? P.wsize : P.sizes[(size_t)w], B - first);
            max_rows = std::max<long long>(max_rows, roff[first + cnt] - roff[first]);
            first += cnt;
        }
        rows_off = slot_bytes;
        slot_bytes = al(slot_bytes + (size_t)max_rows * m.n_lc * 96);
    }
    const int n_slots = (int)std::min<long long>(CM_STAGE_SLOTS, P.n_waves);
    const size_t need = mask_bytes + (size_t)n_slots * slot_bytes;
    if (need > h->stage_bytes) {
        if (h->stage) cudaFree(h->stage);
        h->stage = nullptr; h->stage_bytes = 0;
        CU(cudaMalloc((void **)&h->stage, need));
        h->stage_bytes = need;
    }
    cudaStream_t st = h->stream;
    // CM_WAVE_TRACE=1 (debug aid): device timeline of the call on stderr - per wave, when its first kernel started, when
    // its last kernel ended and when its results were on the host, in ms from the start of the call
    static const bool trace = getenv("CM_WAVE_TRACE") && atoi(getenv("CM_WAVE_TRACE"));
    std::vector<cudaEvent_t> tev;
    if (trace) {
        tev.resize(1 + 3 * (size_t)P.n_waves);
        for (auto &e : tev) CU(cudaEventCreate(&e));
        host_mark[0] = host_ms();
        CU(cudaEventRecord(tev[0], st));
    }
    if (!ready) CU(cudaMemcpyAsync(h->stage, masks_host, Bn * m.mask_words * 4, cudaMemcpyHostToDevice, st));
    int64_t chunks_seen = 0;
    if (aux_host) CU(cudaMemcpyAsync(h->d_geom, aux_host, Bn * sizeof(int), cudaMemcpyHostToDevice, st));
    if (roff) CU(cudaMemcpyAsync(h->d_rowoff, roff, (Bn + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    char *slots = h->stage + mask_bytes;
    auto before = [&](long long w, long long first, int count, cudaStream_t ls, cm_out_dev &od) -> int {
        if (ready) {
            const int64_t need = (first + count - 1) / chunk + 1;           // chunks this wave reads
            for (; chunks_seen < need; ++chunks_seen) {
                int32_t r;
                while ((r = __atomic_load_n(const_cast<const int32_t *>(ready) + chunks_seen, __ATOMIC_ACQUIRE)) == 0)
                    std::this_thread::sleep_for(std::chrono::microseconds(20));
                if (r < 0) return fail(CM_ERR_STATE, "cm_solve_many_host_chunks: cancelled by the producer");
            }
            CU(cudaMemcpyAsync(h->stage + (size_t)first * m.mask_words * 4, masks_host + (size_t)first * m.mask_words,
                               (size_t)count * m.mask_words * 4, cudaMemcpyHostToDevice, ls));
        }
        const int s = (int)(w % n_slots);
        if (w >= n_slots) CU(cudaStreamWaitEvent(ls, h->slot_free[s], 0));      // the slot's previous copies are done
        if (trace) CU(cudaEventRecord(tev[1 + 3 * (size_t)w], ls));
        // the kernels index their outputs by the subset's position in the batch: bias the slot pointers
        char *base = slots + (size_t)s * slot_bytes;
        auto bp = [&](int i) -> char * { return it[i].host ? base + it[i].off - (size_t)first * it[i].stride : nullptr; };
        od.flag = (unsigned char *)bp(0); od.status = (unsigned char *)bp(1); od.n_free = (int *)bp(2); od.n_fixed = (int *)bp(3);
        od.dof_map = (int *)bp(4); od.node_exists = (unsigned char *)bp(5); od.U = (double *)bp(6); od.Rf = (double *)bp(7);
        od.eR = (double *)bp(8); od.compliance = (double *)bp(9); od.max_trans = (double *)bp(10); od.max_rot = (double *)bp(11);
        od.min_pivot = (double *)bp(12); od.profile_flops = (double *)bp(13); od.sigma_max = (double *)bp(14);
        od.max_abs_sigma = (double *)bp(15); od.profile_entries = (double *)bp(16); od.envelope_tiles = (double *)bp(17);
        od.Rf_sup = (double *)bp(18);
        if (roff) { od.eR_rows = (double *)(base + rows_off); od.eR_row_offsets = h->d_rowoff; od.row_bias = roff[first]; }
        return CM_OK;
    };
    // The waves of one lane finish in order, the waves of different lanes do not (wave 1 may be done 3 ms before wave 0):
    // with ONE copy stream the copies would leave in enqueue order, each behind the kernels of every earlier wave.
    static const bool copy_lanes = !(getenv("CM_COPY_LANES") && atoi(getenv("CM_COPY_LANES")) == 0);
    auto copy_stream_of = [&](long long w) { return (P.lanes && copy_lanes) ? h->copy_lane[w % P.NL] : h->copy_stream; };
    auto after = [&](long long w, long long first, int count, cudaStream_t ls) -> int {
        const int s = (int)(w % n_slots);
        char *base = slots + (size_t)s * slot_bytes;
        cudaStream_t cs = copy_stream_of(w);
        CU(cudaEventRecord(h->slot_done[s], ls));
        if (trace) CU(cudaEventRecord(tev[2 + 3 * (size_t)w], ls));
        CU(cudaStreamWaitEvent(cs, h->slot_done[s], 0));
        for (int i = 0; i < N_OUT; ++i)
            if (it[i].host)
                CU(cudaMemcpyAsync(static_cast<char *>(it[i].host) + (size_t)first * it[i].stride, base + it[i].off,
                                   (size_t)count * it[i].stride, cudaMemcpyDeviceToHost, cs));
        if (roff && roff[first + count] > roff[first])
            CU(cudaMemcpyAsync(o->eR_rows + (size_t)(roff[first] - roff[0]) * m.n_lc * 12, base + rows_off,
                               (size_t)(roff[first + count] - roff[first]) * m.n_lc * 96, cudaMemcpyDeviceToHost, cs));
        CU(cudaEventRecord(h->slot_free[s], cs));
        if (trace) CU(cudaEventRecord(tev[3 + 3 * (size_t)w], cs));
        return CM_OK;
    };
    cm_out_dev od0{};
    int rc = run_waves(h, reinterpret_cast<const uint32_t *>(h->stage), geom_host ? h->d_geom : nullptr,
                       trows_host ? h->d_geom : nullptr, trows_host ? group : 0, B, od0, h->wsbuf, P, st, before, after);
    if (rc != CM_OK) { cudaDeviceSynchronize(); return rc; }
    host_mark[1] = host_ms();
    if (P.lanes && copy_lanes) {
        for (int i = 0; i < P.NL; ++i) {
            CU(cudaEventRecord(h->copy_lane_ev[i], h->copy_lane[i]));
            CU(cudaStreamWaitEvent(st, h->copy_lane_ev[i], 0));
        }
    } else {
        CU(cudaEventRecord(h->copy_ev, h->copy_stream));
        CU(cudaStreamWaitEvent(st, h->copy_ev, 0));
    }
    CU(cudaStreamSynchronize(st));
    host_mark[2] = host_ms();
    if (trace) {
        fprintf(stderr, "[cm host] validation + staging set-up until the first event %.3f ms, all waves enqueued at %.3f, synchronised at %.3f\n",
                host_mark[0], host_mark[1], host_mark[2]);
        long long first = 0;
        for (long long w = 0; w < P.n_waves; ++w) {
            const long long cnt = std::min<long long>(P.sizes.empty() ? P.wsize : P.sizes[(size_t)w], B - first);
            float a = 0, b = 0, c = 0;
            cudaEventElapsedTime(&a, tev[0], tev[1 + 3 * (size_t)w]);
            cudaEventElapsedTime(&b, tev[0], tev[2 + 3 * (size_t)w]);
            cudaEventElapsedTime(&c, tev[0], tev[3 + 3 * (size_t)w]);
            fprintf(stderr, "[cm wave %2lld lane %lld] %5lld subsets  start %7.2f  kernels done %7.2f  on host %7.2f ms\n", w, w % P.NL, cnt, a, b, c);
            first += cnt;
        }
        for (auto &e : tev) cudaEventDestroy(e);
    }
    return CM_OK;
}

extern "C" {

int cm_solve_many_host_chunks(cm_handle *h, const uint32_t *masks_host, int64_t B, const cm_solve_out_host *o,
                              int64_t chunk, const volatile int32_t *ready) {
    if (!h || !o || !masks_host) return fail(CM_ERR_INVALID, "cm_solve_many_host_chunks: null argument");
    if (chunk <= 0) return fail(CM_ERR_INVALID, "cm_solve_many_host_chunks: chunk must be positive");
    if (!ready) return solve_host(h, masks_host, nullptr, nullptr, 0, B, o);
    return solve_host(h, masks_host, nullptr, nullptr, 0, B, o, chunk, ready);
}

// One process, several GPUs: the batch is range-partitioned over the handles (one per device, same
// model and load cases), one host thread per device drives its share through cm_solve_many_host, and
// every thread's copies land at the subsets' final positions in the caller's arrays - that IS the
// host gather of the partitioned solve; there is no collective.
int cm_solve_many_host_multi(cm_handle *const *hs, int32_t n_handles, const uint32_t *masks_host, int64_t B,
                             const cm_solve_out_host *o) {
    if (!hs || n_handles < 1 || !o || !masks_host) return fail(CM_ERR_INVALID, "cm_solve_many_host_multi: null argument");
    for (int g = 0; g < n_handles; ++g) {
        if (!hs[g]) return fail(CM_ERR_INVALID, "cm_solve_many_host_multi: null handle");
        if (hs[g]->m.nE != hs[0]->m.nE || hs[g]->m.nV != hs[0]->m.nV || hs[g]->m.n_lc != hs[0]->m.n_lc)
            return fail(CM_ERR_INVALID, "cm_solve_many_host_multi: handles describe different models or load cases");
        for (int k = 0; k < g; ++k)
            if (hs[k] == hs[g] || hs[k]->device == hs[g]->device)
                return fail(CM_ERR_INVALID, "cm_solve_many_host_multi: one handle per device");
    }
    if (B <= 0) return CM_OK;
    if (n_handles == 1) return cm_solve_many_host(hs[0], masks_host, B, o);
    std::vector<int> rcs(n_handles, CM_OK);
    std::vector<std::string> msgs(n_handles);
    std::vector<std::thread> th;
    const int64_t base = B / n_handles, extra = B % n_handles;
    for (int g = 0; g < n_handles; ++g) {
        const int64_t lo = g * base + std::min<int64_t>(g, extra), n = base + (g < extra ? 1 : 0);
        if (n == 0) continue;
        th.emplace_back([=, &rcs, &msgs]() {
            cm_solve_out_host c = *o;
            out_advance(hs[g], &c, lo);
            rcs[g] = cm_solve_many_host(hs[g], masks_host + (size_t)lo * hs[g]->m.mask_words, n, &c);
            if (rcs[g] != CM_OK) msgs[g] = g_err;        // g_err is thread-local: carry the message to the caller's thread
        });
    }
    for (auto &t : th) t.join();
    for (int g = 0; g < n_handles; ++g)
        if (rcs[g] != CM_OK) return fail(rcs[g], "device " + std::to_string(hs[g]->device) + ": " + msgs[g]);
    return CM_OK;
}

int64_t cm_kernel_launches(const cm_handle *h) { return h ? h->launches : 0; }
int cm_reset_counters(cm_handle *h) {
    if (!h) return fail(CM_ERR_INVALID, "cm_reset_counters: null handle");
    h->launches = 0;
    h->stage_ms[0] = h->stage_ms[1] = h->stage_ms[2] = 0;
    return CM_OK;
}
int cm_enable_timing(cm_handle *h, int32_t on) {
    if (!h) return fail(CM_ERR_INVALID, "cm_enable_timing: null handle");
    h->timing = on;
    return CM_OK;
}
int cm_last_stage_ms(const cm_handle *h, double *ms) {
    if (!h || !ms) return fail(CM_ERR_INVALID, "cm_last_stage_ms: null argument");
    for (int i = 0; i < 3; ++i) ms[i] = h->stage_ms[i];
    return CM_OK;
}

int cm_last_wave_count(const cm_handle *h) { return h ? h->last_waves : 0; }

int cm_element_row_offsets(const uint32_t *masks_host, int64_t B, int32_t mask_words, int64_t *offsets) {
    if (!masks_host || !offsets || B < 0 || mask_words < 1) return fail(CM_ERR_INVALID, "cm_element_row_offsets: bad argument");
    int64_t run = 0;
    for (int64_t b = 0; b < B; ++b) {
        offsets[b] = run;
        run += popcount_words(masks_host + (size_t)b * mask_words, mask_words);
    }
    offsets[B] = run;
    return CM_OK;
}

int cm_element_kernel_dev(uint64_t xyz_u, uint64_t xyz_v, uint64_t props, uint64_t cr, int64_t n, uint64_t Kg,
                          uint64_t R3, uint64_t stream) {
    if (n <= 0) return CM_OK;
    if (!xyz_u || !xyz_v || !props || !cr || !Kg || !R3) return fail(CM_ERR_INVALID, "cm_element_kernel_dev: null buffer");
    CU(cm_launch_element_generic((const double *)xyz_u, (const double *)xyz_v, (const double *)props,
                                 (const double *)cr, n, (double *)Kg, (double *)R3, (cudaStream_t)stream));
    return CM_OK;
}

}  // extern "C"
