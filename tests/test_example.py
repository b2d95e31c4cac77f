"""Caller side of the README example (megapose6d_b200/example.py): directory layout, JSON formats, end-to-end run."""
import json

import numpy as np
import pytest
import torch
from PIL import Image

from megapose6d_b200 import example, procedural


def _write_obj(path, mesh, scale_to_mm=1000.0):
    lines = ["mtllib material.mtl", "usemtl m0"]
    lines += ["v %.6f %.6f %.6f" % tuple(v * scale_to_mm) for v in mesh.vertices]
    if mesh.uv is not None:
        lines += ["vt %.6f %.6f" % tuple(t) for t in mesh.uv]
    lines += ["vn %.6f %.6f %.6f" % tuple(n) for n in mesh.vertex_normals]
    for f in mesh.faces:
        if mesh.uv is not None:
            lines.append("f " + " ".join(f"{i + 1}/{i + 1}/{i + 1}" for i in f))
        else:
            lines.append("f " + " ".join(f"{i + 1}//{i + 1}" for i in f))
    path.write_text("\n".join(lines) + "\n")
    if mesh.texture is not None:
        Image.fromarray(mesh.texture).save(path.parent / "texture.png")
        (path.parent / "material.mtl").write_text("newmtl m0\nmap_Kd texture.png\n")


def _make_example_dir(root, rgb, dense=False):
    # `dense`: enough vertices for the pipeline's 2000-point subsets (the reference samples without replacement too)
    box = (procedural.textured_sphere(seed=3) if dense else procedural.textured_box(seed=3)).with_defaults()
    (root / "meshes" / "box").mkdir(parents=True)
    _write_obj(root / "meshes" / "box" / "box.obj", box)
    (root / "inputs").mkdir()
    K = procedural.example_camera()
    (root / "camera_data.json").write_text(json.dumps({"K": K.tolist(), "resolution": [480, 640]}))
    (root / "inputs" / "object_data.json").write_text(json.dumps([{"label": "box", "bbox_modal": [250, 170, 390, 300]}]))
    Image.fromarray(rgb).save(root / "image_rgb.png")
    return K


def test_quaternion_json_round_trip():
    rs = np.random.RandomState(0)
    for _ in range(50):
        q = rs.randn(4)
        q /= np.linalg.norm(q)
        T = example.transform_from_quat_trans(q, rs.randn(3))
        assert np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-12)
        q2 = example.rotmat_to_quat_xyzw(T[:3, :3])
        assert np.allclose(q2, q, atol=1e-9) or np.allclose(q2, -q, atol=1e-9)
        d = example.ObjectData("x", TWO=T, bbox_modal=np.array([1.0, 2, 3, 4])).to_json()
        back = example.ObjectData.from_json(json.loads(json.dumps(d)))
        assert np.allclose(back.TWO, T, atol=1e-9) and np.allclose(back.bbox_modal, [1, 2, 3, 4])


def test_example_directory_readers(tmp_path):
    rgb = np.random.RandomState(1).randint(0, 255, size=(480, 640, 3), dtype=np.uint8)
    K = _make_example_dir(tmp_path, rgb)
    got_rgb, depth, cam = example.load_observation(tmp_path)
    assert depth is None and np.array_equal(got_rgb, rgb) and np.allclose(cam.K, K) and cam.resolution == (480, 640)
    obs = example.load_observation_tensor(tmp_path)
    assert obs.images.shape == (1, 3, 480, 640) and obs.K.shape == (1, 3, 3)
    ds = example.make_object_dataset(tmp_path)
    assert [o.label for o in ds.list_objects] == ["box"] and ds[0].mesh_units == "mm" and abs(ds[0].scale - 0.001) < 1e-12
    det = example.make_detections_from_object_data(example.load_object_data(tmp_path / "inputs" / "object_data.json"))
    assert det.infos["label"].tolist() == ["box"] and det.bboxes.tolist() == [[250.0, 170.0, 390.0, 300.0]]
    # the mesh keeps its texture through the mesh database (millimetres -> metres)
    from megapose6d_b200.meshes import MeshDataBase
    db = MeshDataBase.from_object_ds(ds)
    m = db.meshes["box"]
    assert m.texture is not None and m.uv.shape[0] == m.vertices.shape[0] == 24
    assert abs(np.abs(m.vertices).max() * ds[0].scale - 0.05) < 1e-6


@pytest.mark.gpu
def test_example_runs_end_to_end(tmp_path):
    from megapose6d_b200 import load_model
    from tests import helpers

    rgb = (np.random.RandomState(2).rand(480, 640, 3) * 255).astype(np.uint8)
    _make_example_dir(tmp_path, rgb, dense=True)
    models = tmp_path / "models"
    load_model.write_run(models, "coarse-rgb-906902141", helpers.make_state_dict(helpers.COARSE_CFG, 1))
    load_model.write_run(models, "refiner-rgb-653307694", helpers.make_state_dict(helpers.REFINER_CFG, 2))
    out = example.run_inference(tmp_path, "megapose-1.0-RGB", models_root=models)
    assert len(out) == 1 and torch.isfinite(out.poses).all()
    saved = json.loads((tmp_path / "outputs" / "object_data.json").read_text())
    assert saved[0]["label"] == "box"
    T = example.transform_from_quat_trans(*saved[0]["TWO"])
    assert np.allclose(T, out.poses[0].cpu().double().numpy(), atol=1e-5)
    assert 0.1 < T[2, 3] < 3.0  # in front of the camera, metres
