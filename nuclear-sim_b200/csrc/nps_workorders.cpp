// Work-order table: the host-side bookkeeping of the automatic maintenance loop, native.
//
// Restates what ColumnarAutoMaintenance (nuclear-sim_b200/maintenance.py) does per step with numpy, which in turn restates
//   StateManager._check_maintenance_thresholds          simulator/state/state_manager.py:1307-1369  (violations of one
//                                                        component in one step become ONE event)
//   AutoMaintenanceSystem._create_automatic_work_order  systems/maintenance/auto_maintenance.py:398-466 (known action,
//                                                        24-"hour" dedupe stamp compared in minutes (sic), no active order
//                                                        for the same (component, action), priority delay)
//   AutoMaintenanceSystem.update / _execute_work_order  auto_maintenance.py:200-236, 504-580 (due orders in creation
//                                                        order; at HEAD at most one per plant per update)
// for N plants at once.  Pure host code: plain arrays in, plain arrays out (C ABI in include/nps_b200.h); the decision for
// an event with several violations (MaintenanceOrchestrator, rare) stays with the caller, which patches the group's
// action before nps_wo_issue.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/nps_b200.h"

namespace {

struct Order {
    int64_t plant, comp, act, prio, sub, seq, serial;     // serial: global creation counter (creation order across plants)
    double created, planned;
};
struct Stamp { int64_t key; double t; };                 // key = component * 4096 + action

// A plant holds a handful of stamps and pending orders at a time: the first N live inside the plant's own record (the
// table is walked in plant order, so a step's bookkeeping streams through memory), more spill to the heap.
template <typename T, int N>
struct SmallVec {
    T inl[N];
    std::vector<T>* more = nullptr;
    uint32_t n = 0;
    SmallVec() = default;
    SmallVec(const SmallVec&) = delete;
    SmallVec& operator=(const SmallVec&) = delete;
    SmallVec(SmallVec&& o) noexcept : more(o.more), n(o.n) { for (uint32_t i = 0; i < n && i < (uint32_t)N; ++i) inl[i] = o.inl[i]; o.more = nullptr; o.n = 0; }
    ~SmallVec() { delete more; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T& operator[](size_t i) { return i < (size_t)N ? inl[i] : (*more)[i - N]; }
    const T& operator[](size_t i) const { return i < (size_t)N ? inl[i] : (*more)[i - N]; }
    T& back() { return (*this)[n - 1]; }
    void push_back(const T& v) {
        if (n < (uint32_t)N) inl[n] = v;
        else { if (!more) more = new std::vector<T>(); more->push_back(v); }
        ++n;
    }
    void pop_back() { if (n > (uint32_t)N) more->pop_back(); --n; }
    void erase_at(size_t i) { for (size_t k = i; k + 1 < n; ++k) (*this)[k] = (*this)[k + 1]; pop_back(); }   // keeps the order
    void clear() { if (more) more->clear(); n = 0; }
};

// everything the table knows about one plant
struct PlantBook {
    SmallVec<Stamp, 6> stamps;       // dedupe stamps: creation time of the last work order per (component, action)
    SmallVec<Order, 3> pending;      // scheduled, not yet executed, in creation order
    int64_t n_created = 0;           // WO-%06d numbering
    bool listed = false;             // in nps_wo_table::busy
};

}  // namespace

struct nps_wo_table {
    int64_t n_plants = 0;
    int n_components = 0, n_rows = 0, max_rules = 0;
    double dedupe_window = 24.0;
    double prio_delay[5] = {0, 0, 0, 0, 0};
    bool head_quirks = true;
    // per threshold row: component, rule table (value > rule_thr[j] -> rule_act[j], first match wins), fallback action,
    // the row's own action, priority, sub-component selector
    std::vector<int64_t> row_comp, rule_act, fallback, row_action, row_prio, row_sub;
    std::vector<double> rule_thr;
    std::vector<PlantBook> book;         // one per plant
    std::vector<int64_t> busy;           // plants that have (or recently had) pending orders
    int64_t n_pending = 0, serial = 0;
    struct Due { int64_t plant; size_t index; int64_t serial; };
    std::vector<Due> last_due;           // handed out by nps_wo_due, until nps_wo_complete
};

extern "C" {

int nps_wo_create(int64_t n_plants, int n_components, int n_rows, int max_rules, const int64_t* row_comp,
                  const double* rule_thr, const int64_t* rule_act, const int64_t* fallback, const int64_t* row_action,
                  const int64_t* row_prio, const int64_t* row_sub, const double* prio_delay_minutes, double dedupe_window,
                  int head_quirks, nps_wo_table** out) {
    if (!out || n_plants <= 0 || n_components <= 0 || n_rows <= 0 || max_rules <= 0) return -1;
    nps_wo_table* t = new nps_wo_table();
    t->n_plants = n_plants; t->n_components = n_components; t->n_rows = n_rows; t->max_rules = max_rules;
    t->row_comp.assign(row_comp, row_comp + n_rows);
    t->rule_thr.assign(rule_thr, rule_thr + (size_t)n_rows * max_rules);
    t->rule_act.assign(rule_act, rule_act + (size_t)n_rows * max_rules);
    t->fallback.assign(fallback, fallback + n_rows);
    t->row_action.assign(row_action, row_action + n_rows);
    t->row_prio.assign(row_prio, row_prio + n_rows);
    t->row_sub.assign(row_sub, row_sub + n_rows);
    for (int i = 0; i < 5; ++i) t->prio_delay[i] = prio_delay_minutes[i];
    t->dedupe_window = dedupe_window;
    t->head_quirks = head_quirks != 0;
    t->book.resize((size_t)n_plants);
    *out = t;
    return 0;
}

void nps_wo_destroy(nps_wo_table* t) { delete t; }

// One step's violations sorted by (plant, row) -> one group per (plant, component).  For a group with ONE violation the
// action follows from the row's rule table; groups with several violations get the first row's lookup and are counted in
// *n_multi (the caller decides those).  Outputs are per group; returns the number of groups, -1 on bad input.
int64_t nps_wo_group(const nps_wo_table* t, int64_t n, const int64_t* plant, const int64_t* row, const double* value,
                     int64_t* g_start, int64_t* g_count, int64_t* g_plant, int64_t* g_comp, int64_t* g_act,
                     int64_t* g_prio, int64_t* g_sub, int64_t* n_multi) {
    if (!t || n < 0) return -1;
    int64_t g = -1, multi = 0;
    int64_t last_plant = -1, last_comp = -1;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t r = row[i];
        if (r < 0 || r >= t->n_rows || plant[i] < 0 || plant[i] >= t->n_plants) return -1;
        const int64_t c = t->row_comp[(size_t)r];
        if (g >= 0 && plant[i] == last_plant && c == last_comp) {
            if (g_count[g]++ == 1) ++multi;
            continue;
        }
        ++g;
        last_plant = plant[i]; last_comp = c;
        g_start[g] = i; g_count[g] = 1; g_plant[g] = plant[i]; g_comp[g] = c;
        int64_t act = t->fallback[(size_t)r];
        const double v = value[i];
        for (int j = t->max_rules - 1; j >= 0; --j)              // first matching rule wins: test the last one first
            if (v > t->rule_thr[(size_t)r * t->max_rules + j]) act = t->rule_act[(size_t)r * t->max_rules + j];
        g_act[g] = act;
        g_prio[g] = t->row_prio[(size_t)r];
        g_sub[g] = (t->row_action[(size_t)r] == act) ? t->row_sub[(size_t)r] : 0;
    }
    if (n_multi) *n_multi = multi;
    return g + 1;
}

// Decided events of one step, in (plant, component) order -> work orders.  act_ok[a]: the action is one the maintenance
// catalogue knows.  Writes, for every created order, the index of its group and its per-plant sequence number; returns
// the number created.
int64_t nps_wo_issue(nps_wo_table* t, double t_minutes, int64_t n_groups, const int64_t* g_plant, const int64_t* g_comp,
                     const int64_t* g_act, const int64_t* g_prio, const int64_t* g_sub, const uint8_t* act_ok,
                     int64_t n_actions, int64_t* out_group, int64_t* out_seq) {
    if (!t || n_groups < 0) return -1;
    int64_t made = 0;
    for (int64_t g = 0; g < n_groups; ++g) {
        if (g + 8 < n_groups && g_plant[g + 8] >= 0 && g_plant[g + 8] < t->n_plants) {   // the books are 48 MB at 131 K plants
            const char* nb = reinterpret_cast<const char*>(&t->book[(size_t)g_plant[g + 8]]);
            __builtin_prefetch(nb); __builtin_prefetch(nb + 64); __builtin_prefetch(nb + 128);
        }
        const int64_t a = g_act[g];
        if (a < 0 || a >= n_actions || !act_ok[a]) continue;
        if (a >= 4096 || g_prio[g] < 0 || g_prio[g] > 4 || g_plant[g] < 0 || g_plant[g] >= t->n_plants) return -1;
        PlantBook& b = t->book[(size_t)g_plant[g]];
        const int64_t k = g_comp[g] * 4096 + a;
        bool blocked = false;
        for (size_t i = 0; i < b.stamps.size();) {
            if (!((t_minutes - b.stamps[i].t) < t->dedupe_window)) {      // expired: the clock only moves forward
                b.stamps[i] = b.stamps.back(); b.stamps.pop_back();
                continue;
            }
            if (b.stamps[i].key == k) blocked = true;                     // minutes against the hours constant, sic
            ++i;
        }
        if (blocked) continue;
        for (size_t i = 0; i < b.pending.size(); ++i)
            if (b.pending[i].comp * 4096 + b.pending[i].act == k) { blocked = true; break; }   // an active order for the same (component, action)
        if (blocked) continue;
        Order o;
        o.plant = g_plant[g]; o.comp = g_comp[g]; o.act = a; o.prio = g_prio[g]; o.sub = g_sub[g];
        o.seq = ++b.n_created;
        o.serial = t->serial++;
        o.created = t_minutes;
        o.planned = t_minutes + t->prio_delay[o.prio];
        b.pending.push_back(o);
        b.stamps.push_back(Stamp{k, t_minutes});
        if (!b.listed) { b.listed = true; t->busy.push_back(o.plant); }
        ++t->n_pending;
        out_group[made] = g; out_seq[made] = o.seq;
        ++made;
    }
    return made;
}

int64_t nps_wo_n_pending(const nps_wo_table* t) { return t ? t->n_pending : -1; }

// The orders update(t) executes: due ones in creation order, at most one per plant when head_quirks (the first due one);
// otherwise all of them, plant by plant.  Fills the order columns (capacity `cap` each) and remembers the selection for
// nps_wo_complete.  Returns the count (or the needed capacity when cap is too small, with nothing remembered).
int64_t nps_wo_due(nps_wo_table* t, double t_minutes, int64_t cap, int64_t* plant, int64_t* comp, int64_t* act,
                   int64_t* prio, int64_t* sub, int64_t* seq, double* created, double* planned) {
    if (!t) return -1;
    std::vector<nps_wo_table::Due> due;
    size_t w = 0;
    for (size_t q = 0; q < t->busy.size(); ++q) {            // plants without pending orders drop out of the busy list
        if (q + 8 < t->busy.size()) {
            const char* nb = reinterpret_cast<const char*>(&t->book[(size_t)t->busy[q + 8]].pending);
            __builtin_prefetch(nb); __builtin_prefetch(nb + 64);
        }
        const int64_t p = t->busy[q];
        PlantBook& b = t->book[(size_t)p];
        if (b.pending.empty()) { b.listed = false; continue; }
        t->busy[w++] = p;
        for (size_t i = 0; i < b.pending.size(); ++i) {
            if (t_minutes >= b.pending[i].planned) {
                due.push_back({p, i, b.pending[i].serial});
                if (t->head_quirks) break;
            }
        }
    }
    t->busy.resize(w);
    if (t->head_quirks)      // creation order over all plants
        std::sort(due.begin(), due.end(), [](const nps_wo_table::Due& a, const nps_wo_table::Due& b) { return a.serial < b.serial; });
    else                     // plant by plant, each plant's orders in creation order
        std::sort(due.begin(), due.end(), [](const nps_wo_table::Due& a, const nps_wo_table::Due& b) {
            return a.plant != b.plant ? a.plant < b.plant : a.serial < b.serial; });
    if ((int64_t)due.size() > cap) return (int64_t)due.size();
    for (size_t j = 0; j < due.size(); ++j) {
        const Order& o = t->book[(size_t)due[j].plant].pending[due[j].index];
        plant[j] = o.plant; comp[j] = o.comp; act[j] = o.act; prio[j] = o.prio; sub[j] = o.sub; seq[j] = o.seq;
        created[j] = o.created; planned[j] = o.planned;
    }
    t->last_due = due;
    return (int64_t)due.size();
}

// The orders handed out by the last nps_wo_due have been executed: they leave the table.
int nps_wo_complete(nps_wo_table* t) {
    if (!t) return -1;
    // several orders of one plant (head_quirks off): erase from the back so the remembered indices stay valid
    std::sort(t->last_due.begin(), t->last_due.end(), [](const nps_wo_table::Due& a, const nps_wo_table::Due& b) {
        return a.plant != b.plant ? a.plant < b.plant : a.index > b.index; });
    for (const auto& d : t->last_due) {
        t->book[(size_t)d.plant].pending.erase_at(d.index);
        --t->n_pending;
    }
    t->last_due.clear();
    return 0;
}

// Episode reset of some plants: their pending orders, stamps and numbering go.
int nps_wo_reset_plants(nps_wo_table* t, const int64_t* plants, int64_t n) {
    if (!t) return -1;
    for (int64_t i = 0; i < n; ++i) {
        if (plants[i] < 0 || plants[i] >= t->n_plants) return -1;
        PlantBook& b = t->book[(size_t)plants[i]];
        t->n_pending -= (int64_t)b.pending.size();
        b.pending.clear(); b.stamps.clear(); b.n_created = 0;
    }
    t->last_due.clear();
    return 0;
}

// Checkpoint support: sizes, export and import of the whole table (pending orders in creation order, stamps by plant).
static std::vector<const Order*> all_pending(const nps_wo_table* t) {
    std::vector<const Order*> v;
    for (const PlantBook& b : t->book) for (size_t i = 0; i < b.pending.size(); ++i) v.push_back(&b.pending[i]);
    std::sort(v.begin(), v.end(), [](const Order* a, const Order* b) { return a->serial < b->serial; });
    return v;
}
int nps_wo_sizes(const nps_wo_table* t, int64_t* n_pending, int64_t* n_stamps) {
    if (!t) return -1;
    int64_t s = 0;
    for (const PlantBook& b : t->book) s += (int64_t)b.stamps.size();
    *n_pending = t->n_pending; *n_stamps = s;
    return 0;
}
int nps_wo_export(const nps_wo_table* t, int64_t* pend_cols /*[6][n_pending]*/, double* pend_times /*[2][n_pending]*/,
                  int64_t* stamp_keys /*[2][n_stamps]: plant, component * 4096 + action*/, double* stamp_times, int64_t* n_created) {
    if (!t) return -1;
    const std::vector<const Order*> v = all_pending(t);
    const size_t P = v.size();
    for (size_t i = 0; i < P; ++i) {
        const Order& o = *v[i];
        pend_cols[0 * P + i] = o.plant; pend_cols[1 * P + i] = o.comp; pend_cols[2 * P + i] = o.act;
        pend_cols[3 * P + i] = o.prio; pend_cols[4 * P + i] = o.sub; pend_cols[5 * P + i] = o.seq;
        pend_times[0 * P + i] = o.created; pend_times[1 * P + i] = o.planned;
    }
    size_t S = 0;
    for (const PlantBook& b : t->book) S += b.stamps.size();
    size_t k = 0;
    for (size_t p = 0; p < t->book.size(); ++p)
        for (size_t i = 0; i < t->book[p].stamps.size(); ++i) {
            const Stamp& s = t->book[p].stamps[i];
            stamp_keys[k] = (int64_t)p; stamp_keys[S + k] = s.key; stamp_times[k] = s.t; ++k;
        }
    for (size_t p = 0; p < t->book.size(); ++p) n_created[p] = t->book[p].n_created;
    return 0;
}
int nps_wo_import(nps_wo_table* t, int64_t n_pending, const int64_t* pend_cols, const double* pend_times, int64_t n_stamps,
                  const int64_t* stamp_keys, const double* stamp_times, const int64_t* n_created) {
    if (!t || n_pending < 0 || n_stamps < 0) return -1;
    for (PlantBook& b : t->book) { b.pending.clear(); b.stamps.clear(); b.listed = false; }
    t->busy.clear(); t->last_due.clear(); t->n_pending = 0; t->serial = 0;
    const size_t P = (size_t)n_pending;
    for (size_t i = 0; i < P; ++i) {
        Order o;
        o.plant = pend_cols[0 * P + i]; o.comp = pend_cols[1 * P + i]; o.act = pend_cols[2 * P + i];
        o.prio = pend_cols[3 * P + i]; o.sub = pend_cols[4 * P + i]; o.seq = pend_cols[5 * P + i];
        o.created = pend_times[0 * P + i]; o.planned = pend_times[1 * P + i];
        if (o.plant < 0 || o.plant >= t->n_plants) return -1;
        o.serial = t->serial++;
        PlantBook& b = t->book[(size_t)o.plant];
        b.pending.push_back(o);
        if (!b.listed) { b.listed = true; t->busy.push_back(o.plant); }
        ++t->n_pending;
    }
    const size_t S = (size_t)n_stamps;
    for (size_t i = 0; i < S; ++i) {
        if (stamp_keys[i] < 0 || stamp_keys[i] >= t->n_plants) return -1;
        t->book[(size_t)stamp_keys[i]].stamps.push_back(Stamp{stamp_keys[S + i], stamp_times[i]});
    }
    for (size_t p = 0; p < t->book.size(); ++p) t->book[p].n_created = n_created[p];
    return 0;
}

}  // extern "C"
