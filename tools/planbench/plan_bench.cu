// CPU-side harness (measurement aid): times plan_add_view on dumped views without a GPU.
#include "../../pedfac_b200/csrc/pf_engine.cu"
#include <chrono>
struct V { std::vector<int> a[9]; int n_cc; std::vector<uint8_t> founder, selfing; pf_graph_view v; };
int main(int argc, char **argv)
{
    FILE *f = fopen("/tmp/planbench/views.bin", "rb");
    int ns, ci, cm; fread(&ns, 4, 1, f); fread(&ci, 4, 1, f); fread(&cm, 4, 1, f);
    std::vector<V> vs(ns);
    for (auto &x : vs) {
        fread(&x.n_cc, 4, 1, f);
        for (int k = 0; k < 9; k++) { int n; fread(&n, 4, 1, f); x.a[k].resize(n); fread(x.a[k].data(), 4, n, f); }
        x.founder.assign(x.a[2].begin(), x.a[2].end()); x.selfing.assign(x.a[6].begin(), x.a[6].end());
        pf_graph_view &v = x.v;
        v.n_indiv = (int)x.a[0].size(); v.indiv = x.a[0].data(); v.indiv_cc = x.a[1].data(); v.indiv_founder = x.founder.data();
        v.n_marr = (int)x.a[3].size(); v.marr = x.a[3].data(); v.marr_pa = x.a[4].data(); v.marr_ma = x.a[5].data();
        v.marr_selfing = x.selfing.data(); v.marr_kid_begin = x.a[7].data(); v.marr_kids = x.a[8].data(); v.n_cc = x.n_cc;
    }
    pf_engine *e = new pf_engine();
    e->cfg.n_loci = 2000; e->cfg.locus_begin = 0; e->cfg.locus_end = 2000; e->cfg.cap_indiv = ci; e->cfg.cap_marr = cm; e->cfg.n_chains = 1; e->cfg.device = 0;
    e->Sl = 2000; e->Sp = 2048;
    e->loc.assign(ci, -1); e->h_kind.assign(ci, EMIS_DEC16); e->h_eptr.assign(ci, (void *)0x7f0000000000ull); e->h_denom.assign(ci, 10000); e->h_rcp.assign(ci, 1e-4);
    const int reps = argc > 1 ? atoi(argv[1]) : 200;
    static Plan plan;
    double best = 1e9;
    for (int r = 0; r < reps; r++) {
        auto t0 = std::chrono::steady_clock::now();
        for (auto &x : vs) {
            plan.reset(); plan.target_tasks_per_group = 256;
            int max_up = 0;
            int rc = plan_add_view(e, &x.v, 0, 0, plan, max_up, false, SWEEP_LATENCY);
            if (rc) { printf("rc %d %s\n", rc, pf_last_error()); return 1; }
        }
        double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / vs.size();
        best = std::min(best, us);
    }
    {   // the second-generation planner (emit2) on the same views
        static Plan2 p2; static Plan dummy;
        double best2 = 1e9;
        for (int r = 0; r < reps; r++) {
            auto t0 = std::chrono::steady_clock::now();
            for (auto &x : vs) {
                p2.reset(); p2.target_tasks_per_group = 256; p2.format = 2;
                int max_up = 0;
                int rc = plan_add_view(e, &x.v, 0, 0, dummy, max_up, false, SWEEP_LATENCY, nullptr, &p2);
                if (rc) { printf("rc %d %s\n", rc, pf_last_error()); return 1; }
            }
            best2 = std::min(best2, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / vs.size());
        }
        printf("plan_add_view + emit2: %.1f us per view (best of %d), last plan: %zu tasks %zu pool words %zu groups\n", best2, reps, p2.tasks.size(), p2.pool.size(), p2.groups.size());
    }
    e->host_prof = true; e->n_prof_calls = 1000;
    static Plan2 q2; static Plan dummy2;
    for (int r = 0; r < 100; r++) for (auto &x : vs) { q2.reset(); q2.target_tasks_per_group = 256; q2.format = 2; int mu = 0; plan_add_view(e, &x.v, 0, 0, dummy2, mu, false, SWEEP_LATENCY, nullptr, &q2); }
    printf("v2 phases us: adjacency %.1f traversals %.1f kid lists %.1f rows+levels %.1f records %.1f\n", 1e6*e->t_ph[0]/(100*vs.size()), 1e6*e->t_ph[1]/(100*vs.size()), 1e6*e->t_ph[2]/(100*vs.size()), 1e6*e->t_ph[3]/(100*vs.size()), 1e6*e->t_ph[4]/(100*vs.size()));
    printf("plan_add_view: %.1f us per view (best of %d), last plan: %zu tasks %zu pool ints %zu emis\n", best, reps, plan.tasks.size(), plan.pool.size(), plan.emis.size());
    return 0;
}
