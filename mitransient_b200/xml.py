"""Mitsuba 3 XML scene files -> scene dictionaries (``mi.load_file`` for the subset this build supports).

The reference ships its example scenes as XML (examples/transient/cornell-box/cbox_diffuse.xml,
cbox_volumetric.xml, examples/transient-nlos/nlos_Z.xml): ``<default>`` substitution (cbox_diffuse.xml:2-4),
plugins as nested elements, typed properties, ``<transform>`` blocks and ``<ref>``s.  The result is handed to
:func:`mitransient_b200.load_dict`, so the supported plugin set is the same.
"""
from __future__ import annotations

import os
import re
import xml.etree.ElementTree as ET
from typing import Any, Dict

import numpy as np

from .scene import load_dict
from .transforms import ScalarTransform4f

_PLUGIN_TAGS = {"integrator", "sensor", "sampler", "film", "rfilter", "bsdf", "shape", "emitter", "medium", "phase",
                "texture", "spectrum"}


def _floats(text: str):
    return [float(x) for x in re.split(r"[\s,]+", text.strip()) if x]


def _vec3(el, default=0.0):
    if "value" in el.attrib:
        v = _floats(el.attrib["value"])
        return v * 3 if len(v) == 1 else v
    return [float(el.attrib.get(k, default)) for k in ("x", "y", "z")]


def _transform(el) -> ScalarTransform4f:
    m = np.eye(4)
    for op in el:      # operations apply in file order: each one is multiplied from the left
        T = ScalarTransform4f()
        if op.tag == "translate":
            t = T.translate(_vec3(op))
        elif op.tag == "scale":
            t = T.scale(_vec3(op, 1.0))
        elif op.tag == "rotate":
            t = T.rotate(_vec3(op), float(op.attrib["angle"]))
        elif op.tag == "lookat":
            t = T.look_at(_floats(op.attrib["origin"]), _floats(op.attrib["target"]),
                          _floats(op.attrib.get("up", "0, 1, 0")))
        elif op.tag == "matrix":
            v = _floats(op.attrib["value"])
            t = ScalarTransform4f(np.array(v, dtype=np.float64).reshape(4, 4) if len(v) == 16 else
                                  np.block([[np.array(v).reshape(3, 3), np.zeros((3, 1))], [np.zeros((1, 3)), 1.0]]))
        else:
            raise ValueError(f"unsupported transform operation <{op.tag}>")
        m = t.matrix @ m
    return ScalarTransform4f(m)


def _substitute(root, defaults: Dict[str, str]):
    for el in root.iter():
        for k, v in list(el.attrib.items()):
            if "$" in v:
                for name, val in defaults.items():
                    v = v.replace("$" + name, str(val))
                if "$" in v:
                    raise ValueError(f"unresolved parameter in attribute {k}={el.attrib[k]!r}")
                el.attrib[k] = v


def _plugin(el, counter) -> Dict[str, Any]:
    d: Dict[str, Any] = {"type": el.attrib["type"]}
    for ch in el:
        name = ch.attrib.get("name")
        if ch.tag in ("float",):
            d[name] = float(ch.attrib["value"])
        elif ch.tag == "integer":
            d[name] = int(ch.attrib["value"])
        elif ch.tag == "boolean":
            d[name] = ch.attrib["value"].strip().lower() == "true"
        elif ch.tag == "string":
            d[name] = ch.attrib["value"]
        elif ch.tag in ("rgb", "spectrum") and "type" not in ch.attrib:
            v = _floats(ch.attrib["value"])
            d[name] = {"type": "rgb", "value": v * 3 if len(v) == 1 else v}
        elif ch.tag in ("point", "vector"):
            d[name] = _vec3(ch)
        elif ch.tag == "transform":
            d[name or "to_world"] = _transform(ch)
        elif ch.tag == "ref":
            counter[0] += 1
            d[name or f"_ref{counter[0]}"] = {"type": "ref", "id": ch.attrib["id"]}
        elif ch.tag in _PLUGIN_TAGS:
            counter[0] += 1
            key = name or ch.attrib.get("id") or {"film": "film", "sampler": "sampler", "rfilter": "rfilter",
                                                   "phase": "phase"}.get(ch.tag, f"_{ch.tag}{counter[0]}")
            d[key] = _plugin(ch, counter)
        elif ch.tag in ("default", "alias", "include", "path"):
            continue
        else:
            raise ValueError(f"unsupported XML element <{ch.tag}> in <{el.tag} type={el.attrib.get('type')!r}>")
    return d


def xml_to_dict(path: str, **params) -> Dict[str, Any]:
    """Parse a scene file into the dictionary ``load_dict`` understands.  ``params`` override ``<default>``s."""
    root = ET.parse(path).getroot()
    if root.tag != "scene":
        raise ValueError("the root element must be <scene>")
    defaults = {el.attrib["name"]: el.attrib["value"] for el in root.iter("default")}
    defaults.update({k: v for k, v in params.items()})
    _substitute(root, defaults)
    scene: Dict[str, Any] = {"type": "scene"}
    counter = [0]
    for el in root:
        if el.tag == "default":
            continue
        if el.tag not in _PLUGIN_TAGS:
            raise ValueError(f"unsupported top-level element <{el.tag}>")
        counter[0] += 1
        key = el.attrib.get("id") or f"{el.tag}_{counter[0]}"
        scene[key] = _plugin(el, counter)
    return scene


def load_file(path: str, **params):
    """``mi.load_file(path, **params)``: parse the XML, resolve mesh file names relative to it, build the scene."""
    return load_dict(xml_to_dict(path, **params), base_dir=os.path.dirname(os.path.abspath(path)))


# ------------------------------------------------------------------------------------ dictionary -> XML
_TAG_OF_TYPE = {
    **{t: "integrator" for t in ("transient_path", "transient_nlos_path", "transient_prbvolpath", "path", "volpath")},
    **{t: "sensor" for t in ("perspective", "nlos_capture_meter")},
    **{t: "sampler" for t in ("independent",)},
    **{t: "film" for t in ("transient_hdr_film", "phasor_hdr_film", "hdrfilm")},
    **{t: "rfilter" for t in ("box", "gaussian", "tent")},
    **{t: "bsdf" for t in ("diffuse", "conductor", "dielectric", "twosided", "null", "plastic", "roughplastic",
                           "roughconductor", "bumpmap")},
    **{t: "shape" for t in ("rectangle", "cube", "obj", "ply", "trimesh")},
    **{t: "emitter" for t in ("area", "projector", "angulararea")},
    **{t: "medium" for t in ("homogeneous",)},
    **{t: "phase" for t in ("hg", "isotropic")},
}


def _emit(parent, name, val, mesh_dir, mesh_prefix, counter):
    if isinstance(val, dict) and "type" in val:
        t = val["type"]
        if t == "ref":
            ET.SubElement(parent, "ref", {"id": str(val["id"]), **({"name": name} if name and not name.startswith("_ref") else {})})
            return
        if t == "rgb":
            v = list(val["value"]) if not isinstance(val["value"], (int, float)) else [val["value"]] * 3
            ET.SubElement(parent, "rgb", {"name": name, "value": ", ".join(repr(float(x)) for x in v)})
            return
        tag = _TAG_OF_TYPE.get(t)
        if tag is None:
            raise ValueError(f"dict_to_xml: unknown plugin type {t!r}")
        attrib = {"type": "obj" if t == "trimesh" else t}
        if parent.tag == "scene":
            attrib["id"] = name
        elif name and not name.startswith("_") and tag not in ("film", "sampler", "rfilter", "phase"):
            attrib["name"] = name
        el = ET.SubElement(parent, tag, attrib)
        for k, v in val.items():
            if k in ("type", "vertices", "faces"):
                continue
            _emit(el, k, v, mesh_dir, mesh_prefix, counter)
        if t == "trimesh":                       # in-memory mesh -> Wavefront OBJ next to the scene file
            counter[0] += 1
            fn = f"{mesh_prefix}_{counter[0]}_{name}.obj"
            v = np.asarray(val["vertices"], dtype=np.float64).reshape(-1, 3)
            f = np.asarray(val["faces"], dtype=np.int64).reshape(-1, 3)
            with open(os.path.join(mesh_dir, fn), "w") as fh:
                for p in v:
                    fh.write("v " + " ".join(repr(float(x)) for x in p) + "\n")
                for tri in f:
                    fh.write("f " + " ".join(str(int(i) + 1) for i in tri) + "\n")
            ET.SubElement(el, "string", {"name": "filename", "value": fn})
        return
    if isinstance(val, ScalarTransform4f):
        tr = ET.SubElement(parent, "transform", {"name": name})
        ET.SubElement(tr, "matrix", {"value": " ".join(repr(float(x)) for x in np.asarray(val.matrix, dtype=np.float64).reshape(-1))})
    elif isinstance(val, bool):
        ET.SubElement(parent, "boolean", {"name": name, "value": "true" if val else "false"})
    elif isinstance(val, (int, np.integer)):
        ET.SubElement(parent, "integer", {"name": name, "value": str(int(val))})
    elif isinstance(val, (float, np.floating)):
        ET.SubElement(parent, "float", {"name": name, "value": repr(float(val))})
    elif isinstance(val, str):
        ET.SubElement(parent, "string", {"name": name, "value": val})
    elif isinstance(val, (list, tuple, np.ndarray)) and len(val) == 3:
        ET.SubElement(parent, "vector", {"name": name, "x": repr(float(val[0])), "y": repr(float(val[1])), "z": repr(float(val[2]))})
    else:
        raise ValueError(f"dict_to_xml: cannot write property {name!r} = {val!r}")


def dict_to_xml(scene: Dict[str, Any], path: str) -> str:
    """Write a scene dictionary as a Mitsuba 3 XML file (``trimesh`` shapes become OBJ files next to it) -- the inverse of
    :func:`xml_to_dict` for the supported subset: ``load_file(dict_to_xml(d, path))`` builds the same scene as
    ``load_dict(d)``.  Transforms are written as ``<matrix>`` blocks, floats with full precision."""
    root = ET.Element("scene", {"version": "3.0.0"})
    mesh_dir = os.path.dirname(os.path.abspath(path))
    prefix = os.path.splitext(os.path.basename(path))[0]
    counter = [0]
    for key, val in scene.items():
        if key == "type":
            continue
        _emit(root, key, val, mesh_dir, prefix, counter)
    ET.indent(root)
    ET.ElementTree(root).write(path, encoding="utf-8", xml_declaration=False)
    return path
