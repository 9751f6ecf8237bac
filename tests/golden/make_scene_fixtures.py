"""Scene fixtures for the GPU box (where /root/reference does not exist): the reference's own example scene files,
read HERE through `xml_to_dict` (defaults substituted, OBJ meshes loaded), and written back with OUR exporter
(`dict_to_xml`: one XML + generated OBJ files per scene) under tests/golden/scenes/.

    python tests/golden/make_scene_fixtures.py        (build container only)

`tests/test_reference_scenes.py` checks, wherever the reference tree is present, that loading the fixture gives the
same flat scene descriptor as loading the reference's file; `tests/test_gpu_fullsize.py` renders the fixtures on the
GPU through `load_file` against the oracle.  The files hold scene DATA (a 38-triangle Cornell box, a 6-triangle letter
Z and their plugin parameters), no code."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mitransient_b200 as mitr  # noqa: E402
from mitransient_b200.scene import load_obj, load_ply  # noqa: E402
from mitransient_b200.xml import dict_to_xml, xml_to_dict  # noqa: E402

EXAMPLES = "/root/reference/examples"
OUT = os.path.join(ROOT, "tests", "golden", "scenes")
SCENES = {   # fixture name -> (reference file, <default> overrides)
    "cbox_diffuse": ("transient/cornell-box/cbox_diffuse.xml", {"res": 64, "spp": 16}),
    "cbox_volumetric": ("transient/cornell-box/cbox_volumetric.xml", {"res": 64, "spp": 16}),
    "nlos_Z": ("transient-nlos/nlos_Z.xml", {}),
}


def inline_meshes(d, base_dir):
    """obj / ply shapes -> in-memory trimesh (vertices, faces) so that dict_to_xml writes its own OBJ files."""
    for key, val in d.items():
        if isinstance(val, dict) and val.get("type") in ("obj", "ply"):
            fn = val["filename"] if os.path.isabs(val["filename"]) else os.path.join(base_dir, val["filename"])
            tris = load_ply(fn) if val["type"] == "ply" else load_obj(fn)
            verts = tris.reshape(-1, 3)
            mesh = {k: v for k, v in val.items() if k not in ("type", "filename")}
            mesh.update({"type": "trimesh", "vertices": verts, "faces": np.arange(len(verts)).reshape(-1, 3)})
            d[key] = mesh
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (rel, params) in SCENES.items():
        src = os.path.join(EXAMPLES, rel)
        d = inline_meshes(xml_to_dict(src, **params), os.path.dirname(src))
        path = dict_to_xml(d, os.path.join(OUT, name + ".xml"))
        a = mitr.load_file(src, **params)
        b = mitr.load_file(path)
        da, db = a.describe(a.sensors()[0]), b.describe(b.sensors()[0])
        assert da.n_prims == db.n_prims and bytes(da.film) == bytes(db.film) and bytes(da.sensor) == bytes(db.sensor), name
        print(name, "->", os.path.relpath(path, ROOT), da.n_prims, "primitives")


if __name__ == "__main__":
    main()
