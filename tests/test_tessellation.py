"""Known-answer tests of the surface tessellation used by the case set-up
(/root/reference/source/tit/geom/tessellation.test.cpp:16-215, restated): the
Python helper of titsolver_b200.cases and the C++ facade (include/tit_b200/sph.hpp)
must reproduce the reference's vertex and face numbering."""
import os
import subprocess

import numpy as np
import pytest

from titsolver_b200 import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TET_FACES = [[0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]]

# (vertices, d_max) -> expected (vertices, faces), tessellation.test.cpp:64-215
GOLDEN_3D = {
    "uniform": (
        [[0, 0, 0], [2, 0, 0], [0, 2, 0], [0, 0, 2]],
        [[0, 0, 0], [2, 0, 0], [0, 2, 0], [0, 0, 2], [1, 1, 0], [1, 0, 1], [0, 1, 1]],
        [[2, 4, 0], [4, 1, 0], [1, 5, 0], [5, 3, 0], [3, 6, 0], [6, 2, 0], [1, 4, 5], [4, 2, 6], [5, 6, 3], [4, 6, 5]],
    ),
    "y_longer": (
        [[0, 0, 0], [2, 0, 0], [0, 4, 0], [0, 0, 2]],
        [[0, 0, 0], [2, 0, 0], [0, 4, 0], [0, 0, 2], [0, 2, 0], [1, 2, 0], [1, 0, 1], [0, 2, 1], [1, 1, 0], [1.5, 1, 0], [0.5, 3, 0],
         [0, 1, 1.5], [0, 1, 0.5], [0, 3, 0.5], [1, 1, 0.5], [0.5, 1, 1]],
        [[4, 8, 0], [8, 1, 0], [5, 9, 4], [9, 8, 4], [9, 1, 8], [2, 10, 4], [10, 5, 4], [1, 6, 0], [6, 3, 0], [3, 11, 0], [11, 12, 0],
         [11, 7, 12], [0, 12, 4], [12, 7, 4], [7, 13, 4], [13, 2, 4], [1, 9, 6], [9, 14, 6], [9, 5, 14], [5, 10, 7], [10, 13, 7],
         [10, 2, 13], [6, 15, 3], [15, 11, 3], [15, 7, 11], [7, 15, 5], [15, 14, 5], [15, 6, 14]],
    ),
    "z_longer": (
        [[0, 0, 0], [2, 0, 0], [0, 2, 0], [0, 0, 4]],
        [[0, 0, 0], [2, 0, 0], [0, 2, 0], [0, 0, 4], [1, 1, 0], [1, 0, 2], [0, 0, 2], [0, 1, 2], [1.5, 0, 1], [0.5, 0, 1], [0.5, 0, 3],
         [0, 1, 1], [0, 1.5, 1], [0, 0.5, 3], [1, 0.5, 1], [0.5, 1, 1]],
        [[2, 4, 0], [4, 1, 0], [1, 8, 0], [8, 9, 0], [8, 5, 9], [0, 9, 6], [9, 5, 6], [5, 10, 6], [10, 3, 6], [6, 11, 0], [11, 2, 0],
         [7, 12, 6], [12, 11, 6], [12, 2, 11], [3, 13, 6], [13, 7, 6], [4, 14, 1], [14, 8, 1], [14, 5, 8], [2, 12, 4], [12, 15, 4],
         [12, 7, 15], [7, 13, 5], [13, 10, 5], [13, 3, 10], [5, 14, 7], [14, 15, 7], [14, 4, 15]],
    ),
}


def test_tessellate_2d_known_answer():
    """tessellation.test.cpp:17-62."""
    v, f = cases.tessellate_2d(np.array([[0.0, 0.0], [3.0, 0.0], [0.0, 4.0]]), np.array([[0, 1], [1, 2], [2, 0]], np.uint64), 1.0)
    want_v = [[0, 0], [3, 0], [0, 4], [1, 0], [2, 0], [2.4, 0.8], [1.8, 1.6], [1.2, 2.4], [0.6, 3.2], [0, 3], [0, 2], [0, 1]]
    want_f = [[0, 3], [3, 4], [4, 1], [1, 5], [5, 6], [6, 7], [7, 8], [8, 2], [2, 9], [9, 10], [10, 11], [11, 0]]
    assert np.allclose(v, want_v, rtol=0, atol=1e-14) and f.tolist() == want_f


@pytest.mark.parametrize("name", sorted(GOLDEN_3D))
def test_tessellate_3d_known_answers(name):
    verts, want_v, want_f = GOLDEN_3D[name]
    v, f = cases.tessellate_3d(np.array(verts, float), np.array(TET_FACES), 2.0)
    assert np.allclose(v, want_v, rtol=0, atol=1e-14)
    assert f.tolist() == want_f


def test_tessellate_3d_bounds_every_edge_and_keeps_the_surface_closed():
    verts, faces = cases._box_wall_mesh((1.0, 0.7, 0.4), (1, 1, 1))
    d_max = 0.11
    v, f = cases.tessellate_3d(verts, faces, d_max)
    f = f.astype(np.int64)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    assert np.linalg.norm(v[e[:, 0]] - v[e[:, 1]], axis=1).max() <= d_max * (1 + 1e-12)
    # closed and consistently oriented: every directed edge has exactly one opposite partner
    fwd = {(int(a), int(b)) for a, b in e}
    assert len(fwd) == len(e) and all((b, a) in fwd for a, b in fwd)
    # the area is that of the box
    area = 0.5 * np.linalg.norm(np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]]), axis=1).sum()
    assert abs(area - 2 * (1.0 * 0.7 + 1.0 * 0.4 + 0.7 * 0.4)) < 1e-12


CPP = r"""
#include <cstdio>
#include "tit_b200/sph.hpp"
using namespace tit;
int main() {
  const double cases[3][3] = {{2, 2, 2}, {2, 4, 2}, {2, 2, 4}};
  for (const auto& c : cases) {
    geom::Surface<Vec<double, 3>> s;
    s.append_vert({0.0, 0.0, 0.0}); s.append_vert({c[0], 0.0, 0.0}); s.append_vert({0.0, c[1], 0.0}); s.append_vert({0.0, 0.0, c[2]});
    s.append_face({0, 2, 1}); s.append_face({0, 1, 3}); s.append_face({0, 3, 2}); s.append_face({1, 2, 3});
    const auto t = geom::tessellate(s, 2.0);
    std::printf("V %zu F %zu\n", t.num_verts(), t.num_faces());
    for (const auto& v : t.verts()) std::printf("v %.17g %.17g %.17g\n", v[0], v[1], v[2]);
    for (const auto& f : t.face_verts()) std::printf("f %zu %zu %zu\n", f[0], f[1], f[2]);
  }
}
"""


def test_facade_tessellate_3d_matches(tmp_path):
    """The C++ facade's geom::tessellate (3-D) gives the reference's numbering too."""
    src = tmp_path / "tess.cpp"
    src.write_text(CPP)
    exe = tmp_path / "tess"
    lib_dir = os.path.join(ROOT, "titsolver_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++20", "-O1", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L" + lib_dir, "-ltitgpu",
                           "-Wl,-rpath," + lib_dir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    blocks, cur = [], None
    for line in out:
        tag, *vals = line.split()
        if tag == "V":
            cur = ([], [])
            blocks.append(cur)
        elif tag == "v":
            cur[0].append([float(x) for x in vals])
        else:
            cur[1].append([int(x) for x in vals])
    for name, (v, f) in zip(("uniform", "y_longer", "z_longer"), blocks):
        _, want_v, want_f = GOLDEN_3D[name]
        assert np.allclose(v, want_v, rtol=0, atol=1e-14) and f == want_f, name
