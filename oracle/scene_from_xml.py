"""Parse the scene the reference's generate_xml_content() emits back into arrays —
TEST INFRASTRUCTURE.  Proves "same centres, radii and camera" (BASELINE.json north_star):
what the kernels are handed equals what Mitsuba would have been handed.

Grammar = XMLTemplates of example_renderer.py:9-74 / traj_ball_renderer.py:7-77 /
traj_b0.py:10-60 (HEAD sensor block, BALL_SEGMENT, TAIL floor + light)."""
import re

import numpy as np

_NUM = r"[-+0-9.eE]+"


def _floats(s):
    return [float(x) for x in s.split(",")]


def parse_scene(xml):
    out = {}
    m = re.search(r'<sensor type="perspective">(.*?)</sensor>', xml, re.S)
    sensor = m.group(1)
    look = re.search(r'<lookat origin="([^"]+)" target="([^"]+)" up="([^"]+)"/>', sensor)
    out["origin"], out["target"], out["up"] = (_floats(look.group(k)) for k in (1, 2, 3))
    out["fov"] = float(re.search(r'name="fov" value="(%s)"' % _NUM, sensor).group(1))
    out["near_clip"] = float(re.search(r'name="nearClip" value="(%s)"' % _NUM, sensor).group(1))
    out["far_clip"] = float(re.search(r'name="farClip" value="(%s)"' % _NUM, sensor).group(1))
    out["width"] = int(re.search(r'name="width" value="(\d+)"', sensor).group(1))
    out["height"] = int(re.search(r'name="height" value="(\d+)"', sensor).group(1))
    out["spp"] = int(re.search(r'name="sampleCount" value="(\d+)"', sensor).group(1))

    spheres = re.findall(
        r'<shape type="sphere">\s*<float name="radius" value="(%s)"/>\s*<transform name="toWorld">\s*'
        r'<translate x="(%s)" y="(%s)" z="(%s)"/>\s*</transform>\s*<bsdf type="diffuse">\s*'
        r'<rgb name="reflectance" value="(%s),(%s),(%s)"/>' % ((_NUM,) * 7), xml)
    arr = np.array(spheres, dtype=np.float64).reshape(-1, 7)
    # centres were printed from f32 values with repr-exact '{}' formatting: the round trip is exact
    out["radius"] = arr[:, 0].astype(np.float32)
    out["centers"] = arr[:, 1:4].astype(np.float32)
    out["reflectance"] = arr[:, 4:7].astype(np.float32)

    rects = re.findall(r'<shape type="rectangle">(.*?)</shape>', xml, re.S)
    for body in rects:
        if "emitter" in body:
            sc = re.search(r'<scale x="(%s)" y="(%s)" z="(%s)"/>' % ((_NUM,) * 3), body)
            lk = re.search(r'<lookat origin="([^"]+)" target="([^"]+)" up="([^"]+)"/>', body)
            out["light_half"] = float(sc.group(1))
            out["light_z"] = _floats(lk.group(1))[2]
            out["radiance"] = _floats(re.search(r'name="radiance" value="([^"]+)"', body).group(1))[0]
        else:
            sc = re.search(r'<scale x="(%s)" y="(%s)" z="(%s)"/>' % ((_NUM,) * 3), body)
            tr = re.search(r'<translate x="(%s)" y="(%s)" z="(%s)"/>' % ((_NUM,) * 3), body)
            sx, sy = float(sc.group(1)), float(sc.group(2))
            tx, ty, tz = (float(tr.group(k)) for k in (1, 2, 3))
            # Mitsuba applies the listed transforms in order: scale the unit [-1,1]^2 rectangle, then translate
            out["floor_z"] = tz
            out["floor_min"] = (tx - sx, ty - sy)
            out["floor_max"] = (tx + sx, ty + sy)
    return out
