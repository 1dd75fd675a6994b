"""Shared helpers for the parity tests: build the same synthetic scene on the oracle side
(oracle.host, literal JS restatement) and on the product side (package host mirror)."""
from __future__ import annotations

import importlib
import os

import numpy as np

import synth
from oracle import host as OH
from oracle import refcl as OR

REF_ROOT = "/root/reference"
HAVE_REF = os.path.isdir(REF_ROOT)


def product():
    return importlib.import_module("2015-raytracing_b200")


def make_scene_pair(tmpdir, cols, rows, mesh_uv=(24, 12), **scene_kw):
    """(oracle_scene, product_scene) for the same synthetic XML + mesh."""
    rt = product()
    mesh = synth.synth_mesh(*mesh_uv)
    path = synth.write_scene(tmpdir, **scene_kw)
    tri_dir = os.path.join(str(tmpdir), "tri")
    os.makedirs(tri_dir, exist_ok=True)
    synth.mesh_to_json_file(mesh, os.path.join(tri_dir, "synth.json"))
    o_scene = OH.loadScene(path, cols, rows)
    p_scene = rt.loadScene(path, cols, rows)
    return o_scene, p_scene


def ulp_diff(a, b):
    """Distance in units of last place between two float32 arrays (NaN == NaN, inf == inf)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    d = np.abs(ia - ib)
    d[np.isnan(a) & np.isnan(b)] = 0
    return d
