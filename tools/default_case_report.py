"""Physical sanity report of a dam-break run stored in a `.ttdb` database
(examples/dam_break_2d.cpp or dam_break_3d.cpp, i.e. the reference's default case on the GPU path).

    python tools/default_case_report.py particles.ttdb report.json [thin.ttdb]

Per frame: time (the reference stores t sqrt(g / H), wcsph.cpp:166,187), position of the
surge front (largest x of a fluid particle), highest fluid particle, largest speed,
density range, number of fluid particles outside the tank, non-finite values. The
summary adds the dimensionless time at which the front reaches the far wall — for this
geometry (column 2H x H in a tank 5.366H long) experiments and SPH runs in the
literature put it at about 2.3-2.6 — and the relative change of the fluid's total
energy-like quantities that must stay bounded (mass is constant by construction).
`thin.ttdb` optionally receives every `len/10`-th frame with r, v, rho, p only.
"""
import json
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from titsolver_b200 import ttdb  # noqa: E402

H = 0.6
TANK = (5.366 * H, 4.0 * H, 1.0 * H)


def main():
    src, out = sys.argv[1], sys.argv[2]
    thin = sys.argv[3] if len(sys.argv) > 3 else None
    with ttdb.Storage(src, read_only=True) as s:
        series = s.last_series()
        frames = series.frames()
        rows = []
        nf = None
        thin_db = ttdb.Storage(thin) if thin else None
        thin_series = thin_db.create_series("thinned") if thin_db else None
        stride = max(1, len(frames) // 10)
        for k, frame in enumerate(frames):
            d = {a.name: a.read() for a in frame.arrays() if a.name in ("r", "v", "rho", "p", "gamma", "m")}
            r, v, rho = d["r"], d["v"], d["rho"]
            dim = r.shape[1]
            if nf is None:
                # fluid first, then one fixed particle per wall vertex: the walls do not move
                last = frames[-1].find_array("r").read()
                nf = int(np.nonzero((last != r).any(axis=1))[0].max()) + 1 if (last != r).any() else r.shape[0]
            rf, vf = r[:nf], v[:nf]
            outside = int(((rf < 0.0) | (rf > np.array(TANK[:dim]))).any(axis=1).sum())
            rows.append({
                "frame": k, "time": frame.time, "front_x_over_H": float(rf[:, 0].max() / H), "top_y_over_H": float(rf[:, 1].max() / H),
                "max_speed_over_sqrt_gH": float(np.sqrt((vf**2).sum(1)).max() / np.sqrt(9.81 * H)),
                "rho_min": float(rho[:nf].min()), "rho_max": float(rho[:nf].max()), "outside_tank": outside,
                "non_finite": int(sum((~np.isfinite(x)).sum() for x in d.values())),
                "kinetic_over_initial_potential": float((0.5 * d["m"][:nf] * (vf**2).sum(1)).sum() / (d["m"][:nf] * 9.81 * frames[0].find_array("r").read()[:nf, 1]).sum()),
            })
            if thin_series is not None and (k % stride == 0 or k == len(frames) - 1):
                thin_series.write_particles(frame.time if k else 0.0, {f: d[f] for f in ("r", "v", "rho", "p")}, names=["r", "v", "rho", "p"])
        if thin_db:
            thin_db.close()
    wall_x = TANK[0] / H
    hit = next((row["time"] for row in rows if row["front_x_over_H"] >= wall_x - 0.05), None)
    summary = {
        "source": src, "frames": len(rows), "n_fluid": nf, "n_total": int(r.shape[0]), "dim": int(dim), "t_end": rows[-1]["time"],
        "front_reaches_far_wall_at_t": hit, "max_outside_tank": max(row["outside_tank"] for row in rows),
        "non_finite_total": sum(row["non_finite"] for row in rows), "rho_range": [min(row["rho_min"] for row in rows), max(row["rho_max"] for row in rows)],
        "max_speed_over_sqrt_gH": max(row["max_speed_over_sqrt_gH"] for row in rows),
    }
    with open(out, "w") as f:
        json.dump({"summary": summary, "frames": rows}, f, indent=1)
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
