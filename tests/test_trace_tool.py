"""The Basilisk-trace ingest tool (tests/trace_tool.py, docs/TRACE_SCHEMA.md) proven on a trace the ORACLE wrote: file
format and validation, Sun-table round trip through `orc_set_ephemeris` / `bskenv_set_ephemeris`, both replay back ends.
It pins the tool chain, not parity: no Basilisk output exists in this repository."""
import numpy as np
import pytest

from tests import trace_tool as tt


def _trace(orc, tmp_path, steps=4, seed=3, **cfg):
    row = orc.ic_to_row(orc.sample_ic_dict(np.random.RandomState(seed)))
    acts = np.random.RandomState(seed + 1).randint(0, 3, steps)
    path = str(tmp_path / "trace.npz")
    tt.record_oracle_trace(row, acts, path, **cfg)
    return path


def test_oracle_written_trace_replays_through_the_oracle(orc, tmp_path):
    path = _trace(orc, tmp_path, steps=4)
    trace = tt.load_trace(path)
    assert trace["sun_r"].shape == (5, 3) and trace["obs"].shape == (4, 5)
    rep = tt.compare(trace, tt.run_oracle(trace))
    # the replay sees the Sun only through the table built from the recorded nodes: reproduces the analytic run to rounding
    assert rep["ok"], rep
    assert rep["r_BN_N"]["worst"] < 1e-13 and rep["sigma_BN"]["worst"] < 1e-12
    assert tt.main([path, "--backend", "oracle", "--json", str(tmp_path / "rep.json")]) == 0


def test_a_wrong_trace_is_reported(orc, tmp_path):
    path = _trace(orc, tmp_path, steps=3)
    z = dict(np.load(path))
    z["r_BN_N"] = z["r_BN_N"] * (1.0 + 5e-9)              # 35 mm: above the 1e-9 tolerance
    z["sun_r"] = z["sun_r"]
    bad = str(tmp_path / "bad.npz")
    np.savez(bad, **z)
    rep = tt.compare(tt.load_trace(bad), tt.run_oracle(tt.load_trace(bad)))
    assert not rep["ok"] and not rep["r_BN_N"]["ok"] and rep["v_BN_N"]["ok"]
    assert tt.main([bad]) == 1
    del z["sigma_BR"]
    np.savez(bad, **z)
    with pytest.raises(tt.TraceError):
        tt.load_trace(bad)


def test_moved_sun_nodes_change_the_replay(orc, tmp_path):
    """The recorded Sun really is what the replay uses: turning it by 0.5 rad changes the third-body pull and the panel /
    eclipse geometry."""
    path = _trace(orc, tmp_path, steps=3, seed=5)
    z = dict(np.load(path))
    c, s = np.cos(0.5), np.sin(0.5)
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    z["sun_r"] = z["sun_r"] @ R.T; z["sun_v"] = z["sun_v"] @ R.T
    moved = str(tmp_path / "moved.npz")
    np.savez(moved, **z)
    tr = tt.load_trace(moved)
    rep = tt.compare(tr, tt.run_oracle(tr))
    assert not rep["ok"] and not rep["r_BN_N"]["ok"], rep


@pytest.mark.gpu
def test_oracle_written_trace_replays_through_the_cuda_path(bsk, orc, tmp_path):
    path = _trace(orc, tmp_path, steps=6, seed=9)
    assert tt.main([path, "--backend", "both"]) == 0
