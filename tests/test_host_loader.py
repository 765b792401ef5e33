"""The C++ scene ingest (bidirectional-pathtracing_b200/host/dae_scene.cpp, via libbdpt_host.so) against the
flat scenes the REFERENCE'S OWN loader produced (tests/golden/scenes/*.fscene.gz, dumped by
oracle/ref_harness.cpp from ColladaParser + Application::load + get_static_scene + build_accel).
The comparison is byte for byte: positions, half-edge vertex normals, primitive order, materials,
lights and the placed camera."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_scene

DAE_DIRS = ["/root/reference/dae/sky", os.path.join(ROOT, "oracle", "_ref", "scenes")]
CASES = {  # fixture name -> .dae file name(s) to look for
    "CBspheres_lambertian": ["CBspheres_lambertian.dae"],
    "CBspheres": ["CBspheres.dae"],
    "CBgems": ["CBgems.dae"],
    "CBempty": ["CBempty.dae"],
    "CBbunny": ["CBbunny.dae"],
    "CBcoil": ["CBcoil.dae"],
    "CBspheres_lambertian_occluded": ["../ben/CBspheres_lambertian.dae", "CBspheres_lambertian_occluded.dae"],
    "CBbunny_microfacet_cu": ["CBbunny_microfacet_cu.dae"],     # fixture derived from CBbunny's, see conftest.load_scene
}


def find_dae(names):
    for d in DAE_DIRS:
        for n in names:
            p = os.path.normpath(os.path.join(d, n))
            if os.path.exists(p):
                return p
    return None


@pytest.mark.parametrize("name", sorted(CASES))
def test_loader_reproduces_reference_scene_bytes(name):
    from bdpt import host
    dae = find_dae(CASES[name])
    if dae is None:
        pytest.skip("no .dae input on this box (reference tree and oracle/_ref/scenes both absent)")
    got, ref = host.load_dae(dae), load_scene(name)
    assert got.n_prims == ref.n_prims
    for part in ("camera", "materials", "lights", "prims"):
        assert getattr(got, part).tobytes() == getattr(ref, part).tobytes(), part


def test_loader_applies_screen_size():
    from bdpt import host
    dae = find_dae(CASES["CBspheres"])
    if dae is None:
        pytest.skip("no .dae input on this box")
    a = host.load_dae(dae, 480, 360)
    b = load_scene("CBspheres").with_frame(480, 360)
    assert a.camera.tobytes() == b.camera.tobytes()


MINI = """<?xml version="1.0" encoding="utf-8"?>
<!-- a hand-written Y_UP scene: one quad floor (a polygon with 4 vertices), a sphere, a point light -->
<COLLADA xmlns="http://www.collada.org/2005/11/COLLADASchema" version="1.4.1">
  <asset><up_axis>Y_UP</up_axis></asset>
  <library_cameras><camera id="cam" name="cam"><optics><technique_common><perspective>
    <xfov>40</xfov><aspect_ratio>1.3333</aspect_ratio><znear>0.1</znear><zfar>100</zfar>
  </perspective></technique_common></optics></camera></library_cameras>
  <library_lights><light id="L" name="L"><technique_common><point><color>3 2 1</color>
    <constant_attenuation>1</constant_attenuation><linear_attenuation>0</linear_attenuation><quadratic_attenuation>0</quadratic_attenuation>
  </point></technique_common></light></library_lights>
  <library_effects><effect id="fx"><profile_COMMON><technique sid="common"><phong>
    <diffuse><color sid="diffuse">0.25 0.5 0.75 1</color></diffuse></phong></technique></profile_COMMON></effect>
    <effect id="glassfx"><profile_COMMON><technique sid="common"><phong><diffuse><color>1 1 1 1</color></diffuse></phong></technique></profile_COMMON>
      <extra><technique profile="CGL"><glass><transmittance>1 1 1</transmittance><reflectance>1 1 1</reflectance><roughness>0</roughness><ior>1.5</ior></glass></technique></extra></effect>
  </library_effects>
  <library_materials><material id="mat" name="mat"><instance_effect url="#fx"/></material>
    <material id="glass" name="glass"><instance_effect url="#glassfx"/></material></library_materials>
  <library_geometries>
    <geometry id="floor" name="floor"><mesh>
      <source id="floor-pos"><float_array id="floor-pos-array" count="12">-1 0 -1 -1 0 1 1 0 1 1 0 -1</float_array></source>
      <vertices id="floor-vtx"><input semantic="POSITION" source="#floor-pos"/></vertices>
      <polylist material="mat" count="1"><input semantic="VERTEX" source="#floor-vtx" offset="0"/><vcount>4</vcount><p>0 1 2 3</p></polylist>
    </mesh></geometry>
    <geometry id="ball" name="ball"><extra><technique profile="CGL"><sphere><radius>0.25</radius></sphere></technique></extra></geometry>
  </library_geometries>
  <library_visual_scenes><visual_scene id="Scene">
    <node id="n_cam" name="cam"><matrix>1 0 0 0 0 1 0 1 0 0 1 5 0 0 0 1</matrix><instance_camera url="#cam"/></node>
    <node id="n_light" name="light"><matrix>1 0 0 0 0 1 0 2 0 0 1 0 0 0 0 1</matrix><instance_light url="#L"/></node>
    <node id="n_floor" name="floor"><matrix>2 0 0 0 0 1 0 0 0 0 2 0 0 0 0 1</matrix><instance_geometry url="#floor">
      <bind_material><technique_common><instance_material symbol="mat" target="#mat"/></technique_common></bind_material></instance_geometry></node>
    <node id="n_ball" name="ball"><matrix>2 0 0 0.5 0 2 0 0.5 0 0 2 0 0 0 0 1</matrix><instance_geometry url="#ball">
      <bind_material><technique_common><instance_material symbol="glass" target="#glass"/></technique_common></bind_material></instance_geometry></node>
  </visual_scene></library_visual_scenes>
  <scene><instance_visual_scene url="#Scene"/></scene>
</COLLADA>
"""


def test_hand_written_scene(tmp_path):
    from bdpt import flatscene as fs, host
    p = tmp_path / "mini.dae"
    p.write_text(MINI)
    sc = host.load_dae(str(p))
    assert sc.n_prims == 2 and len(sc.materials) == 2 and len(sc.lights) == 1
    tri, sph = sc.prims[0], sc.prims[1]
    # a quad contributes only the triangle (v[n-1], v[0], v[1]) (object.cpp:36-41)
    assert tri["type"] == fs.PRIM_TRIANGLE
    assert np.allclose(tri["data"][:9], [2, 0, -2, -2, 0, -2, -2, 0, 2])
    assert np.allclose(np.abs(tri["data"][9:].reshape(3, 3)), [[0, 1, 0]] * 3)
    assert sph["type"] == fs.PRIM_SPHERE and np.allclose(sph["data"][:4], [0.5, 0.5, 0.0, 0.5])   # radius * |M e_x|
    assert sc.materials[0]["type"] == fs.MAT_DIFFUSE and np.allclose(sc.materials[0]["reflectance"], [0.25, 0.5, 0.75])
    assert sc.materials[1]["type"] == fs.MAT_GLASS and sc.materials[1]["ior"] == np.float32(1.5)
    assert sc.lights[0]["type"] == fs.LIGHT_POINT and np.allclose(sc.lights[0]["position"], [0, 2, 0])
    cam = sc.camera[0]
    # placed looking at the bbox centroid from direction unit(M (0,0,-1,1)) at 1.5 |bbox diagonal|
    target = np.array([0.0, 0.5, 0.0])
    assert np.allclose(cam["target"], target)
    dirc = np.array([0, 1, 4.0]); dirc /= np.linalg.norm(dirc)
    diag = np.linalg.norm([4, 1, 4])
    assert np.allclose(cam["pos"], target + dirc * diag * 1.5, atol=1e-9)
    assert cam["screen_w"] == 800 and cam["screen_h"] == 600


def test_loader_errors(tmp_path):
    from bdpt import host
    with pytest.raises(RuntimeError, match="cannot open"):
        host.load_dae(str(tmp_path / "missing.dae"))
    bad = tmp_path / "bad.dae"
    bad.write_text("<notcollada/>")
    with pytest.raises(RuntimeError, match="not a COLLADA"):
        host.load_dae(str(bad))
    spot = tmp_path / "dir.dae"
    spot.write_text(MINI.replace("<point><color>3 2 1</color>", "<directional><color>3 2 1</color>").replace("</point>", "</directional>"))
    with pytest.raises(RuntimeError, match="light type not supported"):
        host.load_dae(str(spot))
    nocam = tmp_path / "nocam.dae"
    nocam.write_text(MINI.replace('<instance_camera url="#cam"/>', ""))
    with pytest.raises(RuntimeError, match="no camera"):
        host.load_dae(str(nocam))


def read_png(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(raw):
        n, typ = struct.unpack(">I4s", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(typ + body)
        if typ == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert (depth, ctype) == (8, 6)
        elif typ == b"IDAT":
            idat += body
        pos += 12 + n
    px = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 4 * w)
    assert (px[:, 0] == 0).all()
    return px[:, 1:].reshape(h, w, 4)


@pytest.mark.gpu
def test_cli_renders_like_the_library(tmp_path):
    """`pathtracer -s -m -r -f` end to end: .dae -> PNG equals the image the library tone-maps for the same scene."""
    from bdpt import api
    dae = find_dae(CASES["CBspheres"])
    if dae is None:
        pytest.skip("no .dae input on this box")
    exe = os.path.join(ROOT, "bidirectional-pathtracing_b200", "host", "pathtracer")
    out = tmp_path / "cli.png"
    r = subprocess.run([exe, "-t", "8", "-s", "16", "-m", "5", "-r", "160", "120", "-f", str(out), dae], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Rendering... 100%!" in r.stdout and "million rays per second" in r.stdout
    img = read_png(str(out))
    R = api.Renderer(0, ns_aa=16, max_ray_depth=5, seed=0)
    R.set_scene(load_scene("CBspheres")); R.set_frame_size(160, 120); R.render()
    ref = R.tonemap(0)
    assert img.shape == ref.shape == (120, 160, 4)
    assert np.mean(np.abs(img.astype(int) - ref.astype(int)) <= 1) > 0.999
    rate = read_png(str(out)[:-4] + "_rate.png")
    assert (rate[..., 0] == 255).all() and (rate[..., 1:3] == 0).all()


# ------------------------------------------------------------------ `-c <camera settings file>` and `-p <cell>`
def read_png_any(path):
    """RGBA8 array of any PNG (lodepng, which the reference writes with, drops an all-opaque alpha channel)."""
    Image = pytest.importorskip("PIL.Image")
    return np.asarray(Image.open(path).convert("RGBA"))


def write_camera_settings(path, cam, pos=None, hfov=None, vfov=None):
    """A file in Camera::dump_settings' format (reference camera.cpp:152-167) for the given flat-scene camera."""
    c = cam[0]
    c2w = np.array(c["c2w"]).reshape(3, 3).T          # columns -> c2w(row, col)
    p = np.array(c["pos"]) if pos is None else np.array(pos, float)
    r = lambda v: repr(float(v))
    with open(path, "w") as f:
        f.write(f"{r(c['hfov_deg'] if hfov is None else hfov)} {r(c['vfov_deg'] if vfov is None else vfov)} 1.3333 {r(c['nclip'])} {r(c['fclip'])}\n")
        f.write(" ".join(r(v) for v in list(p) + list(c["target"])) + " \n")
        f.write("0 0 5 0.1 100\n")
        f.write(" ".join(r(v) for v in c2w.reshape(-1)) + " \n")
        f.write(f"{r(c['screen_w'])} {r(c['screen_h'])} {r(c['screen_dist'])}\n")
        f.write("4.56 0.0\n")


def test_camera_settings_file_against_the_reference_binary(tmp_path):
    """`-c file` (Camera::load_settings, camera.cpp:169-186) is pinned by the reference ITSELF: pathtracer_ref renders
    CBspheres_lambertian with `-r 48 36 -c file` (a moved camera with its own fovs) on one thread; the fp64 oracle, given
    the camera our parser produces for the same file, reproduces the PNG byte for byte (rng_mode 0 = the reference's
    own random streams).  This also pins the ORDER: the file is applied after -r, its fovs win (main.cpp:174-178)."""
    from bdpt import host
    from oracle import oracle as O
    dae = find_dae(CASES["CBspheres_lambertian"])
    if dae is None or not os.path.exists(O.REF_CLI):
        pytest.skip("needs the reference binary and its .dae inputs (oracle/_ref)")
    sc = host.load_dae(dae)
    W, H, ns, m = 48, 36, 2, 3
    cfg = tmp_path / "cam.txt"
    pos = np.array(sc.camera[0]["pos"]) + np.array([0.0, 0.0, -0.8])     # dolly in along the view axis (c2w unchanged)
    write_camera_settings(cfg, sc.camera, pos=pos, hfov=40.0, vfov=31.0)
    out = tmp_path / "ref.png"
    subprocess.run([O.REF_CLI, "-t", "1", "-s", str(ns), "-m", str(m), "-r", str(W), str(H), "-c", str(cfg), "-f", str(out), dae],
                   check=True, capture_output=True)
    ref = read_png_any(str(out))
    cam = host.apply_camera_settings(host.load_dae(dae, W, H), str(cfg), W, H)
    assert cam.camera[0]["hfov_deg"] == 40.0 and cam.camera[0]["vfov_deg"] == 31.0 and np.allclose(cam.camera[0]["pos"], pos)
    img = O.Oracle(cam, W, H).render(ns, m, rng_mode=0)[0]
    px = (np.clip(np.power(np.maximum(img[::-1], 0) * np.sqrt(2.0), 1 / 2.2), 0, 1) * 255).astype(np.uint8)   # toColor + flip
    assert ref.shape == (H, W, 4)
    assert np.array_equal(ref[..., :3], px), "the oracle with the parsed camera reproduces the reference's -c render"
    # and without -c the image is a different one
    subprocess.run([O.REF_CLI, "-t", "1", "-s", str(ns), "-m", str(m), "-r", str(W), str(H), "-f", str(out), dae], check=True, capture_output=True)
    assert not np.array_equal(read_png_any(str(out))[..., :3], px)


def test_cell_render_against_the_reference_binary(tmp_path):
    """`-p x y dx dy` (render_to_file's cell path, raytraced_renderer.cpp:338-346): the oracle with the same cell
    reproduces the reference's cell PNG byte for byte — only the cell's pixels are sampled, the image is dx x dy."""
    from bdpt import host
    from oracle import oracle as O
    dae = find_dae(CASES["CBspheres_lambertian"])
    if dae is None or not os.path.exists(O.REF_CLI):
        pytest.skip("needs the reference binary and its .dae inputs (oracle/_ref)")
    W, H, ns, m = 48, 36, 2, 3
    x, y, dx, dy = 10, 6, 20, 16
    out = tmp_path / "cell.png"
    subprocess.run([O.REF_CLI, "-t", "1", "-s", str(ns), "-m", str(m), "-r", str(W), str(H), "-p", str(x), str(y), str(dx), str(dy),
                    "-f", str(out), dae], check=True, capture_output=True)
    ref = read_png_any(str(out))
    assert ref.shape == (dy, dx, 4)
    orc = O.Oracle(host.load_dae(dae, W, H), W, H)
    orc.set_cell(x, y, dx, dy)
    img = orc.render(ns, m, rng_mode=0)[0]
    px = (np.clip(np.power(np.maximum(img[::-1], 0) * np.sqrt(2.0), 1 / 2.2), 0, 1) * 255).astype(np.uint8)
    assert np.array_equal(ref[..., :3], px[H - y - dy:H - y, x:x + dx])


@pytest.mark.gpu
def test_cli_camera_file_and_cell(tmp_path):
    """`pathtracer -c file -p x y dx dy`: the CLI's image is the crop of what the library renders for the same camera."""
    from bdpt import api, host
    dae = find_dae(CASES["CBspheres"])
    if dae is None:
        pytest.skip("no .dae input on this box")
    W, H, ns, m = 96, 72, 16, 4
    x, y, dx, dy = 24, 10, 40, 32
    sc = host.load_dae(dae)
    cfg = tmp_path / "cam.txt"
    write_camera_settings(cfg, sc.camera, pos=np.array(sc.camera[0]["pos"]) + np.array([0.3, 0.1, -0.5]), hfov=45.0, vfov=35.0)
    exe = os.path.join(ROOT, "bidirectional-pathtracing_b200", "host", "pathtracer")
    out = tmp_path / "cell.png"
    r = subprocess.run([exe, "-s", str(ns), "-m", str(m), "-r", str(W), str(H), "-c", str(cfg), "-p", str(x), str(y), str(dx), str(dy),
                        "-f", str(out), dae], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Cell job completed" in r.stdout
    img = read_png(str(out))
    assert img.shape == (dy, dx, 4)
    cam = host.apply_camera_settings(host.load_dae(dae, W, H), str(cfg), W, H)
    R = api.Renderer(0, ns_aa=ns, max_ray_depth=m, seed=0)
    R.set_scene(cam); R.set_frame_size(W, H); R.set_cell(x, y, dx, dy); R.render()
    ref = R.tonemap(0)[H - y - dy:H - y, x:x + dx]
    assert np.mean(np.abs(img.astype(int) - ref.astype(int)) <= 1) > 0.999
    rate = read_png(str(out)[:-4] + "_rate.png")
    assert rate.shape == (H, W, 4) and (rate[H - y - 1, x, :3] == [255, 0, 0]).all() and (rate[0, 0, :3] == [0, 0, 255]).all()


# ------------------------------------------------------------------ `-e <environment map>`: the map readers
def write_exr(path, img, compression="none", dtype="half", channels=("B", "G", "R")):
    """Minimal single-part scan-line OpenEXR writer (test input): img (h, w, 3) RGB; compression none / zips / zip."""
    h, w, _ = img.shape
    comp = {"none": 0, "rle": 1, "zips": 2, "zip": 3}[compression]
    ptype, npt = (1, np.float16) if dtype == "half" else (2, np.float32)
    def attr(name, typ, data):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(data)) + data
    chl = b"".join(c.encode() + b"\0" + struct.pack("<iBBBBii", ptype, 0, 0, 0, 0, 1, 1) for c in sorted(channels)) + b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    hdr = (struct.pack("<II", 20000630, 2) + attr("channels", "chlist", chl) + attr("compression", "compression", bytes([comp])) +
           attr("dataWindow", "box2i", box) + attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", b"\0") +
           attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)) + attr("screenWindowCenter", "v2f", struct.pack("<ff", 0, 0)) +
           attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0")
    lines = 16 if comp == 3 else 1
    idx = {"R": 0, "G": 1, "B": 2}
    blocks = []
    for y0 in range(0, h, lines):
        raw = b"".join(img[y, :, idx[c]].astype(npt).tobytes() for y in range(y0, min(h, y0 + lines)) for c in sorted(channels))
        data = raw
        if comp in (2, 3):
            a = np.frombuffer(raw, np.uint8)
            t = np.concatenate([a[0::2], a[1::2]]).astype(np.int32)            # interleave
            t[1:] = (t[1:] - t[:-1] + 128 + 256) % 256                          # predictor
            z = zlib.compress(t.astype(np.uint8).tobytes())
            data = z if len(z) < len(raw) else raw
        blocks.append(struct.pack("<ii", y0, len(data)) + data)
    off = len(hdr) + 8 * len(blocks)
    table = b""
    for b in blocks:
        table += struct.pack("<Q", off); off += len(b)
    open(path, "wb").write(hdr + table + b"".join(blocks))


@pytest.mark.parametrize("compression,dtype", [("none", "float"), ("none", "half"), ("zips", "half"), ("zip", "float"), ("zip", "half")])
def test_exr_reader_against_the_reference_reader(tmp_path, compression, dtype):
    """host/envmap_read.h against the reference's OWN EXR reader (the vendored tinyexr, called like load_exr does,
    main.cpp:40-78 — ref_harness exr) on files an independent writer made: every pixel identical."""
    from bdpt import host
    from oracle import oracle as O
    if not O.have_ref():
        pytest.skip("needs oracle/_ref/ref_harness")
    # 37 rows: a ragged last ZIP block.  Smooth and coarsely quantised so that every block really compresses (the
    # reference's tinyexr has no path for the "stored raw because compression did not help" case of the format)
    yy, xx = np.mgrid[0:37, 0:50]
    img = np.stack([np.round(8 * (1 + np.sin(xx / 7.0)) * (1 + yy / 9.0)) / 4, np.round(xx / 5.0) * 0.5 + 0.25, np.round(16 * np.exp(yy / 12.0)) / 16], -1).astype(np.float32)
    img[0, 0] = [0, 2.0 ** -14, 6e4]
    f = tmp_path / "env.exr"
    write_exr(str(f), img, compression, dtype)
    got = host.load_envmap(str(f))
    out = tmp_path / "ref.bin"
    subprocess.run([O.REF_HARNESS, "exr", str(f), str(out)], check=True, capture_output=True)
    ref = np.fromfile(out, np.float64)
    assert (ref[0], ref[1]) == (50, 37)
    ref = ref[2:].reshape(37, 50, 3).astype(np.float32)
    assert got.shape == (37, 50, 3)
    assert np.array_equal(got, ref)
    want = img.astype(np.float16).astype(np.float32) if dtype == "half" else img
    assert np.array_equal(got, want)


def test_pfm_and_constant_environment_maps(tmp_path):
    from bdpt import host
    img = np.arange(4 * 6 * 3, dtype=np.float32).reshape(4, 6, 3)
    f = tmp_path / "e.pfm"
    open(f, "wb").write(b"PF\n6 4\n-1.0\n" + img[::-1].astype("<f4").tobytes())   # little-endian, bottom row first
    assert np.array_equal(host.load_envmap(str(f)), img)
    c = host.load_envmap("0.5,2,3.5")
    assert c.shape == (2, 1, 3) and np.array_equal(c[0, 0], [0.5, 2, 3.5])
    with pytest.raises(RuntimeError):
        host.load_envmap(str(tmp_path / "missing.exr"))


@pytest.mark.gpu
def test_cli_environment_map(tmp_path):
    """`pathtracer -e map.exr`: the environment light joins the scene's lights (raytraced_renderer.cpp:117-119) and the
    image equals what the library renders for the same scene and map."""
    from bdpt import api, host, flatscene as fs
    dae = find_dae(CASES["CBspheres"])
    if dae is None:
        pytest.skip("no .dae input on this box")
    W, H, ns, m = 96, 72, 16, 4
    rng = np.random.default_rng(4)
    env = rng.uniform(0.2, 2.0, (8, 16, 3)).astype(np.float32)
    f = tmp_path / "env.exr"
    write_exr(str(f), env, "zip", "float")
    exe = os.path.join(ROOT, "bidirectional-pathtracing_b200", "host", "pathtracer")
    out = tmp_path / "e.png"
    r = subprocess.run([exe, "-s", str(ns), "-m", str(m), "-r", str(W), str(H), "-e", str(f), "-f", str(out), dae], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Loading environment map" in r.stdout
    img = read_png(str(out))
    sc = host.load_dae(dae)
    lights = np.concatenate([sc.lights, np.zeros(1, fs.LIGHT_DT)])
    lights[-1]["type"] = fs.LIGHT_ENV; lights[-1]["radiance"] = 1.0
    R = api.Renderer(0, ns_aa=ns, max_ray_depth=m, seed=0)
    R.set_scene(fs.FlatScene(sc.camera, sc.materials, lights, sc.prims)); R.set_frame_size(W, H)
    R.set_environment_map(host.load_envmap(str(f))); R.render()
    ref = R.tonemap(0)
    assert np.mean(np.abs(img.astype(int) - ref.astype(int)) <= 1) > 0.999
    P = api.Renderer(0, ns_aa=ns, max_ray_depth=m, seed=0); P.set_scene(sc); P.set_frame_size(W, H); P.render()
    assert np.mean(np.abs(P.tonemap(0).astype(int) - ref.astype(int)) <= 1) < 0.9      # the environment does light the scene
