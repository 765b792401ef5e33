"""Builds the documented CBlucy stand-ins from the largest Cornell-box mesh the reference checkout has.

dae/sky/CBlucy.dae is absent from the reference tree.  SURVEY.md §8d proposes refining
CBbunny's mesh to CBlucy's triangle class; this script does that on the .dae text so that BOTH binaries
(the reference's pathtracer and ours) load the same file:

    python scripts/make_standin.py            # -> oracle/_ref/scenes/CBbunny4.dae (114,304-triangle bunny)
    python scripts/make_standin.py 2          # -> oracle/_ref/scenes/CBbunny16.dae (457,216 triangles)

Each level splits every triangle in four.  New edge vertices use Loop's edge rule 3/8 (a + b) + 1/8 (c + d)
on interior edges and the midpoint on boundary edges; old vertices stay where they are.  The mesh stays an
oriented manifold (HalfedgeMesh::build requires it, halfEdgeMesh.cpp:166-175,359-363).  Outputs go only to
oracle/_ref/scenes/ (git-ignored inputs, shipped to the GPU box); nothing derived from the reference's
files is committed.
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/dae/sky/CBbunny.dae"
OUT_DIR = os.path.join(ROOT, "oracle", "_ref", "scenes")


def subdivide(pos, tris):
    edge_id, edge_opp = {}, {}
    for t in tris:
        for k in range(3):
            a, b, c = int(t[k]), int(t[(k + 1) % 3]), int(t[(k + 2) % 3])
            e = (a, b) if a < b else (b, a)
            edge_opp.setdefault(e, []).append(c)
    new_pos = [pos]
    nv = len(pos)
    extra = np.zeros((len(edge_opp), 3))
    for i, (e, opp) in enumerate(edge_opp.items()):
        edge_id[e] = nv + i
        a, b = pos[e[0]], pos[e[1]]
        extra[i] = 0.375 * (a + b) + 0.125 * (pos[opp[0]] + pos[opp[1]]) if len(opp) == 2 else 0.5 * (a + b)
    new_pos.append(extra)
    out = np.zeros((len(tris) * 4, 3), np.int64)
    for i, t in enumerate(tris):
        a, b, c = map(int, t)
        ab = edge_id[(a, b) if a < b else (b, a)]
        bc = edge_id[(b, c) if b < c else (c, b)]
        ca = edge_id[(c, a) if c < a else (a, c)]
        out[4 * i:4 * i + 4] = [[a, ab, ca], [ab, b, bc], [ca, bc, c], [ab, bc, ca]]
    return np.concatenate(new_pos), out


def main():
    levels = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    txt = open(SRC).read()
    g0 = txt.index('<geometry id="Mesh-mesh"')
    g1 = txt.index("</geometry>", g0) + len("</geometry>")
    geo = txt[g0:g1]
    fa = re.search(r'<float_array id="Mesh-mesh-positions-array" count="(\d+)">([^<]*)</float_array>', geo)
    pos = np.array(fa.group(2).split(), np.float64).reshape(-1, 3)
    p = np.array(re.search(r"<p>([^<]*)</p>", geo).group(1).split(), np.int64).reshape(-1, 2)[:, 0].reshape(-1, 3)
    for _ in range(levels):
        pos, p = subdivide(pos, p)
    pos32 = pos.astype(np.float32)
    new_geo = f'''<geometry id="Mesh-mesh" name="Mesh">
      <mesh>
        <source id="Mesh-mesh-positions">
          <float_array id="Mesh-mesh-positions-array" count="{pos32.size}">{" ".join(repr(float(v)) for v in pos32.ravel())}</float_array>
          <technique_common>
            <accessor source="#Mesh-mesh-positions-array" count="{len(pos32)}" stride="3">
              <param name="X" type="float"/><param name="Y" type="float"/><param name="Z" type="float"/>
            </accessor>
          </technique_common>
        </source>
        <vertices id="Mesh-mesh-vertices">
          <input semantic="POSITION" source="#Mesh-mesh-positions"/>
        </vertices>
        <polylist material="Default-material" count="{len(p)}">
          <input semantic="VERTEX" source="#Mesh-mesh-vertices" offset="0"/>
          <vcount>{"3 " * len(p)}</vcount>
          <p>{" ".join(map(str, p.ravel()))}</p>
        </polylist>
      </mesh>
    </geometry>'''
    os.makedirs(OUT_DIR, exist_ok=True)
    name = f"CBbunny{4 ** levels}.dae"
    with open(os.path.join(OUT_DIR, name), "w") as f:
        f.write(txt[:g0] + new_geo + txt[g1:])
    print(name, "vertices", len(pos32), "triangles", len(p))


if __name__ == "__main__":
    main()
