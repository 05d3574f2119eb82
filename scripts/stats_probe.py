import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "little-physics-engine_b200"))
import lpe_bh, bench
for wlname in sys.argv[1:] or ["c2"]:
    wl = bench.WORKLOADS[wlname]
    x, y, vx, vy, m = lpe_bh.workload(wl["kind"], wl["n"], wl["seed"], bench.U)
    bh = lpe_bh.BarnesHut(0)
    bh.set_instrumentation(timing=True, counts=True)
    bh.upload(x, y, vx, vy, m)
    p = lpe_bh.make_params(bench.U, bench.EPS, theta=bench.THETA, dt_kick=bench.DT, dt_drift=bench.DT)
    bh.step(p, 1)
    st = bh.stats()
    n = wl["n"]; nw = (n + 31) // 32
    print(wlname, "acc/body %.1f  lane-visits/body %.1f  warp-visits/warp %.1f  lane efficiency %.3f  nodes/body %.2f" % (
        st["interactions"] / n, st["visits"] / n, st["warp_visits"] / nw, st["visits"] / (32.0 * st["warp_visits"]), st["n_nodes"] / n))
    nw_ = nw; print("   per warp:", dict(zip(["A-clean","A-dirty","O-dirty","M->accept","M->open","M->split","rounds","frontier"], [round(v/nw_,1) for v in st["t2_kinds"]])), "overflow", st["overflow_chunks"])
    bh.close()
