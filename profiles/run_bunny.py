"""One bunny registration (BASELINE.json configs[1]) for launch lists: python profiles/run_bunny.py [level] [variant] [n_runs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
import workloads as W
capi = graft.load_package().capi
level = sys.argv[1] if len(sys.argv) > 1 else "difficult"
variant = sys.argv[2] if len(sys.argv) > 2 else "pt2pt"
src, tgt, T_gt = W.bunny_problem(level, seed=2)
ctx = capi.Context(0)
ctx.set_cloud(capi.SOURCE, src)
ctx.set_cloud(capi.TARGET, tgt)
p = capi.default_params(variant=variant, entry=capi.RUN_SE3_ICP, reuse_features=0, estimated_overlap=1.0, max_num_se3_iterations=10,
                        mse=1e-5, mse_switch_error=5e-5, number_of_nn_for_LRF=90)
for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 1):
    T, st = ctx.run(p)
print("bunny %s se3_%s: %d iterations (%d SE3), %.2f ms total, %.2f ms setup, %d launches, corr stage %.2f ms" %
      (level, variant, st.num_iterations, st.num_pure_se3_iterations, st.time_total_ms, st.time_setup_ms, st.kernel_launches,
       st.time_se3_correspondence_search_ms))
print("queries searched: %d of %d (%.1f %%)" % (st.queries_searched, st.num_iterations * len(src), 100.0 * st.queries_searched / (st.num_iterations * len(src))))
