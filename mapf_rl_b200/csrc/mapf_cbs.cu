// mapf_cbs.cu — the CBS expert of the reference (search.py:58-442) as host code: conflict-based search over space-time A*,
// used offline to mint solvable test instances (test.py:23-79 `create_test`) and expert action scripts (`find_path`).
// irregular, pointer-chasing, a few instances at a time -- it stays on the host (C++, one thread per
// instance); the GPU contributes what it is good at: the instances (mapf_env_reset), the per-agent distance maps used as
// the low-level heuristic (bfs_navi_kernel's int32 output == search.compute_heuristics, search.py:24-55) and the replay
// of the resulting scripts (mapf_env_rollout).  Nothing here launches a kernel.
//
// What is reproduced: the optimum.  CBS returns a collision-free set of paths of minimal SUM OF COSTS (search.py:17-21)
// whatever conflict it splits on, so cost parity with the reference is exact; the paths themselves are one of several
// optimal sets -- the reference picks the conflict and the constrained agent with random.choice (search.py:316,249) and
// stops after 5 s of wall clock (:320), i.e. its own output is not reproducible from run to run.  Here the search is
// deterministic (first conflict of the node, standard two-way split instead of disjoint splitting; ties in the open lists
// broken like the reference's tuples: (f, h, cell) in the low level, search.py:131, (cost, #conflicts, generation order)
// in the high level, :297) and bounded by a node budget and an optional wall-clock limit.
// Same conventions as the reference: a path waits at its goal for ever (get_location, :78-84), an agent may not finish
// before its last constraint (:171-179), nodes at timestep >= max_steps are not expanded (:186-187), vertex and edge
// (swap) conflicts (:223-243), action ids of environment.py:12 in the emitted script.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <queue>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "mapf_common.cuh"

namespace {

struct Grid {
    int L;
    const uint8_t *obst;  // [L*L], 1 = obstacle
    bool free_cell(int x, int y) const { return x >= 0 && y >= 0 && x < L && y < L && obst[x * L + y] == 0; }
};

// negative constraints of one agent: vertex (t, cell) and edge (t, from -> to), t = arrival time
struct Constraint {
    int agent, t, from, to;  // from < 0: vertex constraint on `to`
    bool operator==(const Constraint &o) const { return agent == o.agent && t == o.t && from == o.from && to == o.to; }
};

struct ConstraintTable {
    std::unordered_set<uint64_t> vertex, edge;
    int last_t = 0;
    static uint64_t vkey(int t, int c) { return ((uint64_t)t << 32) | (uint32_t)c; }
    static uint64_t ekey(int t, int a, int b) { return ((uint64_t)t << 40) | ((uint64_t)a << 20) | (uint64_t)b; }
    void build(const std::vector<Constraint> &cs, int agent)
    {
        for (const Constraint &c : cs) {
            if (c.agent != agent) continue;
            if (c.from < 0) vertex.insert(vkey(c.t, c.to));
            else edge.insert(ekey(c.t, c.from, c.to));
            last_t = std::max(last_t, c.t);
        }
    }
    bool blocked(int from, int to, int t) const
    {
        if (vertex.empty() && edge.empty()) return false;
        return vertex.count(vkey(t, to)) != 0 || edge.count(ekey(t, from, to)) != 0;
    }
};

// BFS distances to `goal` over free cells (search.compute_heuristics, search.py:24-55); -1 = unreachable
void bfs_dist(const Grid &g, int goal, std::vector<int32_t> &dist)
{
    const int L = g.L;
    dist.assign((size_t)L * L, -1);
    std::vector<int> q;
    q.reserve((size_t)L * L);
    dist[goal] = 0;
    q.push_back(goal);
    for (size_t h = 0; h < q.size(); ++h) {
        const int c = q[h], x = c / L, y = c % L;
        const int nx[4] = {x - 1, x + 1, x, x}, ny[4] = {y, y, y - 1, y + 1};
        for (int k = 0; k < 4; ++k)
            if (g.free_cell(nx[k], ny[k]) && dist[nx[k] * L + ny[k]] < 0) {
                dist[nx[k] * L + ny[k]] = dist[c] + 1;
                q.push_back(nx[k] * L + ny[k]);
            }
    }
}

// space-time A* (search.py:148-219): the shortest path of one agent that respects its constraints
bool a_star(const Grid &g, int start, int goal, const int32_t *h, const ConstraintTable &tab, int max_steps, std::vector<int> &path)
{
    struct Node {
        int cell, t, parent;
    };
    struct Open {
        int f, h, cell, id;
        bool operator>(const Open &o) const
        {
            if (f != o.f) return f > o.f;
            if (h != o.h) return h > o.h;
            if (cell != o.cell) return cell > o.cell;  // the reference compares the (x, y) tuples: row-major order
            return id > o.id;
        }
    };
    const int L = g.L;
    if (h[start] < 0) return false;
    std::vector<Node> nodes;
    std::priority_queue<Open, std::vector<Open>, std::greater<Open>> open;
    std::unordered_set<uint64_t> seen;  // (cell, t): every successor costs exactly 1, so the first visit is the best one
    nodes.push_back({start, 0, -1});
    open.push({h[start], h[start], start, 0});
    seen.insert(ConstraintTable::vkey(0, start));
    while (!open.empty()) {
        const Open cur = open.top();
        open.pop();
        const Node n = nodes[cur.id];
        if (n.cell == goal && n.t >= tab.last_t) {
            path.clear();
            for (int id = cur.id; id >= 0; id = nodes[id].parent) path.push_back(nodes[id].cell);
            std::reverse(path.begin(), path.end());
            return true;
        }
        if (n.t >= max_steps) continue;
        const int x = n.cell / L, y = n.cell % L;
        // the reference's successor order (search.py:12-14 `move` + stay): (0,-1), (1,0), (0,1), (-1,0), stay
        const int nx[5] = {x, x + 1, x, x - 1, x}, ny[5] = {y - 1, y, y + 1, y, y};
        for (int k = 0; k < 5; ++k) {
            if (!g.free_cell(nx[k], ny[k])) continue;
            const int c = nx[k] * L + ny[k];
            if (h[c] < 0) continue;  // cannot reach the goal from there
            if (tab.blocked(n.cell, c, n.t + 1)) continue;
            if (!seen.insert(ConstraintTable::vkey(n.t + 1, c)).second) continue;
            nodes.push_back({c, n.t + 1, cur.id});
            open.push({n.t + 1 + h[c], h[c], c, (int)nodes.size() - 1});
        }
    }
    return false;
}

inline int at(const std::vector<int> &p, int t) { return t < (int)p.size() ? p[t] : p.back(); }

struct Conflict {
    int a1, a2, t, c1, c2;  // c2 < 0: both at c1 at time t; else a1 moves c1 -> c2 while a2 moves c2 -> c1, arriving at t
};

// first conflict of every pair (search.py:223-259)
void detect(const std::vector<std::vector<int>> &paths, std::vector<Conflict> &out)
{
    out.clear();
    const int n = (int)paths.size();
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
            const int len = (int)std::max(paths[i].size(), paths[j].size());
            for (int t = 0; t < len; ++t) {
                const int a = at(paths[i], t), b = at(paths[j], t);
                if (a == b) {
                    out.push_back({i, j, t, a, -1});
                    break;
                }
                const int a2 = at(paths[i], t + 1), b2 = at(paths[j], t + 1);
                if (a == b2 && b == a2) {
                    out.push_back({i, j, t + 1, a, a2});
                    break;
                }
            }
        }
}

int sum_of_costs(const std::vector<std::vector<int>> &paths)
{
    int s = 0;
    for (const auto &p : paths) s += (int)p.size() - 1;
    return s;
}

struct HighNode {
    int cost;
    std::vector<Constraint> constraints;
    std::vector<std::vector<int>> paths;
    std::vector<Conflict> conflicts;
};

struct CbsResult {
    bool solved = false;
    int cost = -1;
    long long expanded = 0;
    std::vector<std::vector<int>> paths;
};

CbsResult cbs(const Grid &g, int N, const int *starts, const int *goals, const int32_t *dist_in, int max_steps, int time_limit_ms,
              long long node_limit)
{
    CbsResult res;
    const size_t cells = (size_t)g.L * g.L;
    std::vector<std::vector<int32_t>> own;
    std::vector<const int32_t *> h(N);
    if (dist_in) {
        for (int a = 0; a < N; ++a) h[a] = dist_in + (size_t)a * cells;
    } else {
        own.resize(N);
        for (int a = 0; a < N; ++a) {
            bfs_dist(g, goals[a], own[a]);
            h[a] = own[a].data();
        }
    }
    // mapf_env_bfs_navi marks unreachable cells with INT32_MAX, bfs_dist with -1: normalise to "negative = unreachable"
    std::vector<std::vector<int32_t>> norm;
    if (dist_in) {
        norm.resize(N);
        for (int a = 0; a < N; ++a) {
            norm[a].assign(h[a], h[a] + cells);
            for (auto &v : norm[a])
                if (v == MAPF_DIST_UNREACHABLE) v = -1;
            h[a] = norm[a].data();
        }
    }
    const auto t_begin = std::chrono::steady_clock::now();
    auto root = std::make_shared<HighNode>();
    root->paths.resize(N);
    const ConstraintTable empty;
    for (int a = 0; a < N; ++a)
        if (!a_star(g, starts[a], goals[a], h[a], empty, max_steps, root->paths[a])) return res;  // search.py:309
    root->cost = sum_of_costs(root->paths);
    detect(root->paths, root->conflicts);
    struct Entry {
        int cost, nconf;
        long long gen;
        std::shared_ptr<HighNode> node;
        bool operator>(const Entry &o) const
        {
            if (cost != o.cost) return cost > o.cost;
            if (nconf != o.nconf) return nconf > o.nconf;
            return gen > o.gen;
        }
    };
    std::priority_queue<Entry, std::vector<Entry>, std::greater<Entry>> open;
    long long gen = 0;
    open.push({root->cost, (int)root->conflicts.size(), gen++, root});
    while (!open.empty()) {
        auto P = open.top().node;
        open.pop();
        if (P->conflicts.empty()) {
            res.solved = true;
            res.cost = P->cost;
            res.paths = P->paths;
            return res;
        }
        if (++res.expanded > node_limit) return res;
        if (time_limit_ms > 0 &&
            std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t_begin).count() > time_limit_ms)
            return res;  // search.py:320
        const Conflict c = P->conflicts.front();
        // two children: the conflict is forbidden to one of its agents
        for (int side = 0; side < 2; ++side) {
            Constraint k;
            k.agent = side == 0 ? c.a1 : c.a2;
            k.t = c.t;
            if (c.c2 < 0) k.from = -1, k.to = c.c1;
            else if (side == 0) k.from = c.c1, k.to = c.c2;
            else k.from = c.c2, k.to = c.c1;
            auto Q = std::make_shared<HighNode>();
            Q->constraints = P->constraints;
            if (std::find(Q->constraints.begin(), Q->constraints.end(), k) == Q->constraints.end()) Q->constraints.push_back(k);
            Q->paths = P->paths;
            ConstraintTable tab;
            tab.build(Q->constraints, k.agent);
            if (!a_star(g, starts[k.agent], goals[k.agent], h[k.agent], tab, max_steps, Q->paths[k.agent])) continue;
            Q->cost = sum_of_costs(Q->paths);
            detect(Q->paths, Q->conflicts);
            open.push({Q->cost, (int)Q->conflicts.size(), gen++, Q});
        }
    }
    return res;
}

// paths -> the action script of find_path (search.py:396-442): every path padded with waits to the makespan, one row of N
// action ids (environment.py:12) per step
int emit_actions(const CbsResult &r, int L, int N, uint8_t *actions, int max_T)
{
    size_t len = 0;
    for (const auto &p : r.paths) len = std::max(len, p.size());
    const int T = (int)len - 1;
    if (T > max_T) return -2;
    for (int t = 1; t <= T; ++t)
        for (int a = 0; a < N; ++a) {
            const int c0 = at(r.paths[a], t - 1), c1 = at(r.paths[a], t);
            const int dx = c1 / L - c0 / L, dy = c1 % L - c0 % L;
            uint8_t act = 0;
            if (dx == -1) act = 1;
            else if (dx == 1) act = 2;
            else if (dy == -1) act = 3;
            else if (dy == 1) act = 4;
            actions[(size_t)(t - 1) * N + a] = act;
        }
    return T;
}

int solve_one(const uint8_t *map, int L, int N, const uint8_t *starts, const uint8_t *goals, const int32_t *dist, int max_steps,
              int time_limit_ms, long long node_limit, uint8_t *actions, int max_T, int32_t *T_out, int32_t *cost_out,
              int64_t *expanded_out)
{
    Grid g{L, map};
    std::vector<int> s(N), t(N);
    for (int a = 0; a < N; ++a) {
        s[a] = starts[2 * a] * L + starts[2 * a + 1];
        t[a] = goals[2 * a] * L + goals[2 * a + 1];
        if (starts[2 * a] >= L || starts[2 * a + 1] >= L || goals[2 * a] >= L || goals[2 * a + 1] >= L || map[s[a]] || map[t[a]])
            return MAPF_EINVAL;
    }
    const CbsResult r = cbs(g, N, s.data(), t.data(), dist, max_steps, time_limit_ms, node_limit);
    if (expanded_out) *expanded_out = r.expanded;
    *T_out = -1;
    if (cost_out) *cost_out = -1;
    if (!r.solved) return MAPF_OK;
    const int T = emit_actions(r, L, N, actions, max_T);
    if (T == -2) return MAPF_OK;  // longer than the caller's script buffer: reported as unsolved
    *T_out = T;
    if (cost_out) *cost_out = r.cost;
    return MAPF_OK;
}

}  // namespace

extern "C" {

int mapf_cbs_solve(const uint8_t *h_map, int32_t map_length, int32_t num_agents, const uint8_t *h_starts, const uint8_t *h_goals,
                   const int32_t *h_dist, int32_t max_steps, int32_t time_limit_ms, int64_t node_limit, uint8_t *h_actions_out,
                   int32_t max_T, int32_t *T_out, int32_t *cost_out, int64_t *expanded_out)
{
    if (!h_map || !h_starts || !h_goals || !h_actions_out || !T_out || map_length < 2 || map_length > MAPF_MAX_MAP_SIDE ||
        num_agents < 1 || num_agents > MAPF_MAX_AGENTS || max_steps < 1 || max_T < 0 || node_limit < 1) {
        mapf_set_error("mapf_cbs_solve: bad argument");
        return MAPF_EINVAL;
    }
    const int rc = solve_one(h_map, map_length, num_agents, h_starts, h_goals, h_dist, max_steps, time_limit_ms, node_limit,
                             h_actions_out, max_T, T_out, cost_out, expanded_out);
    if (rc != MAPF_OK) mapf_set_error("mapf_cbs_solve: start / goal outside the map or on an obstacle");
    return rc;
}

int mapf_cbs_solve_batch(int32_t n, const uint8_t *h_maps, int32_t map_length, int32_t num_agents, const uint8_t *h_starts,
                         const uint8_t *h_goals, const int32_t *h_dist, int32_t max_steps, int32_t time_limit_ms, int64_t node_limit,
                         uint8_t *h_actions_out, int32_t max_T, int32_t *T_out, int32_t *cost_out, int64_t *expanded_out,
                         int32_t threads)
{
    if (n < 0 || !h_maps || !h_starts || !h_goals || !h_actions_out || !T_out || map_length < 2 || map_length > MAPF_MAX_MAP_SIDE ||
        num_agents < 1 || num_agents > MAPF_MAX_AGENTS || max_steps < 1 || max_T < 0 || node_limit < 1) {
        mapf_set_error("mapf_cbs_solve_batch: bad argument");
        return MAPF_EINVAL;
    }
    const size_t cells = (size_t)map_length * map_length, N = (size_t)num_agents;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, n));
    std::vector<int> rcs((size_t)std::max(n, 1), MAPF_OK);
    auto work = [&](int tid) {
        for (int i = tid; i < n; i += nt)
            rcs[i] = solve_one(h_maps + (size_t)i * cells, map_length, num_agents, h_starts + (size_t)i * N * 2,
                               h_goals + (size_t)i * N * 2, h_dist ? h_dist + (size_t)i * N * cells : nullptr, max_steps, time_limit_ms,
                               node_limit, h_actions_out + (size_t)i * max_T * N, max_T, T_out + i, cost_out ? cost_out + i : nullptr,
                               expanded_out ? expanded_out + i : nullptr);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
    for (int i = 0; i < n; ++i)
        if (rcs[i] != MAPF_OK) {
            mapf_set_error("mapf_cbs_solve_batch: start / goal outside the map or on an obstacle");
            return rcs[i];
        }
    return MAPF_OK;
}

}  // extern "C"
