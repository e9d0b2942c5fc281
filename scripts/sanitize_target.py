"""Small launches of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.package()
text = open(os.path.join(g.ROOT, "tests", "golden", "valeurs")).read()


def run(features, kernel=0, opts=(), hist=0, n=(0, 3, 5000), dump=True):
    cfg = pkg.Configuration.parse(text, features).with_num_events(25000)
    with pkg.Simulator(cfg, kernel) as sim:
        for k, v in opts:
            sim.set_option(k, v)
        if hist:
            sim.histograms_enable(hist)
        accs = sim.simulate_batches(*n)
        m = sim.simulate_merged(*n)
        if dump and "faster-evgen" not in features:
            sim.rng_dump(1, 2400)
        extra = f" histogram entries {sum(sim.histograms_fetch().counts[0])}" if hist else ""
        print(features or "default", kernel, dict(opts), [a.selected_events for a in accs], m.selected_events, extra, flush=True)


# --schedule: only the launch shapes of the dynamic schedule added in sessions 52-64 (taper stages, big units of 16, f32 packed kernel)
if "--schedule" in sys.argv:
    run("", opts=(("unit_batches", 8), ("taper_units", 2), ("tail_singles", 3)), n=(2, 43, 777), dump=False)
    run("", opts=(("unit_batches", 16), ("taper_units", 1), ("tail_singles", 2), ("grid_warps", 2)), n=(0, 47, 10000), dump=False)
    run("f32", opts=(("unit_batches", 4), ("taper_units", 3), ("tail_singles", 2)), n=(1, 31, 1234), dump=False)
    run("", opts=(("taper_units", -1), ("align_units", 0)), n=(0, 9, 5000), dump=False)
    sys.exit(0)
# every generator / precision / seeding through the fused kernels, in-kernel ordered fold included
for features, kernel in [("", 0), ("", 1), ("standard-random", 0), ("f32", 0), ("standard-random,f32", 0), ("multi-threading,faster-threading", 0)]:
    run(features, kernel)
# schedules: several batches per unit (the stream continues), static rounds, a one-warp grid; batches cut into parts (short last batch:
# empty trailing parts)
run("", opts=(("unit_batches", 2), ("grid_warps", 1), ("sched_dynamic", 0)), n=(2, 7, 777))
run("", opts=(("unit_batches", 3), ("sched_dynamic", 1)), n=(2, 7, 777))
for parts in (2, 5, 10):
    run("", opts=(("batch_parts", parts), ("unit_batches", 3), ("grid_warps", 2), ("sched_dynamic", 0)), n=(2, 4, 1234))
run("f32", opts=(("batch_parts", 5),), n=(2, 4, 1234))
# faster-evgen: the stream pipeline (walk -> records -> physics) in one pass and in many small passes with the redo path,
# the round-1 pipeline in both modes, the xoshiro scan in both modes, the host-walk cross-check
run("faster-evgen")
run("faster-evgen,f32")
run("faster-evgen", opts=(("fe_pass_segments", 160), ("fe_seg_rounds", 64), ("fe_warm", 3)), n=(1, 5, 1234))
run("faster-evgen", opts=(("fe_legacy", 1),))
run("faster-evgen,f32", opts=(("fe_split", 1),))
run("faster-evgen,standard-random")
run("faster-evgen,standard-random,f32", opts=(("fe_split", 1),))
run("faster-evgen", opts=(("fe_split", 1), ("fe_host_scan", 1)))
# per-event observables: default generator and faster-evgen (stream pipeline)
run("", hist=200)
run("faster-evgen,no-photon-sorting", hist=64)
