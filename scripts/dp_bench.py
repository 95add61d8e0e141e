import ctypes as C, json, sys, os
sys.path.insert(0, "tests")
import c6lib
L = c6lib.product()
for n_games in (65536, 262144):
    rec = (c6lib.Record * n_games)()
    st = c6lib.DpStats()
    assert L.c6_default_policy_play(50, 0xC6, 0, n_games, 0, 0, rec, None, None, C.byref(st)) == 0, L.c6_last_error()
    secs = st.kernel_ms * 1e-3
    print(json.dumps({"games": n_games, "kernel_ms": st.kernel_ms, "games_per_s": n_games / secs, "rollout_stones_per_s": st.rollout_stones / secs, "scan_per_search": st.children_scanned / st.searches, "sel_per_search": st.select_steps/st.searches}))
