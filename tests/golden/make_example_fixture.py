"""Generates tests/golden/example_scene.npz from the reference's bundled AICON report.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_example_fixture.py

Parsing follows util/io/reader/aicon/AICONReportFileReader.java:133-387 of the reference
(section switches :135-152, regexes :187,:217,:245,:269,:271, "c = -Ck" :329, "fest" => fixed :315,
lines ending in '***' skipped :214, camera created with RADIAL/TANGENTIAL/AFFINITY/DISTANCE models :308)
and the datum choice of example/ExampleReport.java:71-82 (names longer than 3 characters are not datum).
The fixture stores only numbers (no reference source code).
"""
import os
import re
import sys

import numpy as np

SRC = '/root/reference/JAICOV/example/example.htm'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'example_scene.npz')

RE_OBJ = re.compile(r'^\w+\s+[\d\.+-]+\s+[\d\.+-]+\s+[\d\.+-]+\s+[\d\.]+\s+[\d\.]+\s+[\d\.]+\s+\d+\s+\d+')
RE_IMG = re.compile(r'^\w+\s+\d+\s+[\d\.+-]+\s+[\d\.+-]+\s+[\d\.+-]+\s+[\d\.+-]+\s+[\d\.]+\s+[\d\.]+\s+[\d\.]+\s+[\d\.]+\s+[\d\.]+\s+[\d\.]+')
RE_EOXYZ = re.compile(r'^\d+\s+\d+\s+[\d\.+-]+\s+[\d\.+-]+\s+[\d\.+-]+\s+[\d\.]+\s+[\d\.]+\s+[\d\.]+\s+\d+')
RE_EOANG = re.compile(r'^air\s+rad\s+[\d\.+-]+\s+[\d\.+-]+\s+[\d\.+-]+\s+[\d\.]+\s+[\d\.]+\s+[\d\.]+\s+[\d\.]+\s+[\d\.]+')
RE_BAR = re.compile(r'^\w+\s+\w+\s+[\d\.+-]+.+')


def fullmatch(rx, line):
    # Java String.matches == full match; the patterns above end open, so emulate with match + tail check
    return rx.fullmatch(line) is not None or (rx.pattern.endswith('.+') and rx.fullmatch(line) is not None)


def main():
    section = None
    io = {}
    io_fixed = {}
    r0 = None
    images = {}     # id -> dict
    img_order = []
    cur_img = None
    points = {}     # name -> xyz
    pt_order = []
    obs = []        # (name, img, x, y, sx, sy)
    bars = []
    with open(SRC, encoding='latin-1') as f:
        for raw in f:
            line = raw.strip()
            if '#Start' in line or 'zum Anfang' in line:
                section = None
            if 'name="interior_orientations"' in line: section = 'IO'
            if 'name="exterior_orientations"' in line: section = 'EO'
            if 'name="object_points"' in line: section = 'OBJ'
            if 'name="image_coordinates"' in line: section = 'IMG'
            if 'name="distances"' in line: section = 'BAR'
            try:
                if section == 'IO':
                    if ':' not in line:
                        continue
                    cols = re.split(r'[:\s]+', line)
                    if len(cols) != 3:
                        continue
                    typ = cols[0].strip()
                    if typ.endswith('/R0'):
                        r0 = float(cols[2])
                    if r0 is None:
                        continue
                    value = float(cols[1])
                    fixed = re.fullmatch(r'\w+', cols[2].strip()) is not None
                    io[typ] = value
                    io_fixed[typ] = fixed
                elif section == 'EO':
                    if RE_EOXYZ.fullmatch(line):
                        c = line.split()
                        cur_img = int(c[0])
                        if cur_img not in images:
                            images[cur_img] = {'eo': [0.0] * 6}
                            img_order.append(cur_img)
                        images[cur_img]['eo'][0:3] = [float(c[2]), float(c[3]), float(c[4])]
                    elif cur_img is not None and RE_EOANG.fullmatch(line):
                        c = line.split()
                        images[cur_img]['eo'][3:6] = [float(c[2]), float(c[3]), float(c[4])]
                elif section == 'OBJ':
                    if not RE_OBJ.fullmatch(line):
                        continue
                    c = line.split()
                    if len(c) != 9:
                        continue
                    if c[0] not in points:
                        pt_order.append(c[0])
                    points[c[0]] = [float(c[1]), float(c[2]), float(c[3])]
                elif section == 'IMG':
                    if line.endswith('***'):
                        continue
                    if not RE_IMG.fullmatch(line):
                        continue
                    c = line.split()
                    if len(c) != 12:
                        continue
                    name, img = c[0], int(c[1])
                    if name not in points or img not in images:
                        continue
                    obs.append((name, img, float(c[2]), float(c[3]), float(c[6]), float(c[7])))
                elif section == 'BAR':
                    if not RE_BAR.fullmatch(line):
                        continue
                    c = line.split()
                    if len(c) < 7:
                        continue
                    a, b = c[0], c[1]
                    if a not in points or b not in points or a == b:
                        continue
                    bars.append((a, b, float(c[2]), float(c[5])))
            except Exception:
                continue
    # Image.add ignores a second observation of the same object point in one image (camera/Image.java:56-58)
    pidx = {n: i for i, n in enumerate(pt_order)}
    per_img = {i: {} for i in img_order}
    for (name, img, x, y, sx, sy) in obs:
        if pidx[name] not in per_img[img]:
            per_img[img][pidx[name]] = (x, y, sx, sy)
    obj, xy, sig, ptr = [], [], [], [0]
    for i in img_order:
        for p, (x, y, sx, sy) in per_img[i].items():
            obj.append(p); xy.append((x, y)); sig.append((sx, sy))
        ptr.append(len(obj))
    np.savez_compressed(
        OUT,
        r0=r0,
        # x0, y0, c  (c = -Ck, AICONReportFileReader.java:329)
        io_val=np.array([io['Xh'], io['Yh'], -io['Ck']]),
        io_fixed=np.array([io_fixed['Xh'], io_fixed['Yh'], io_fixed['Ck']]),
        # evaluation order AFFINITY(Cx,Cy), TANGENTIAL(Bx,By), RADIAL(A1..A3), DISTANCE(none in this report)
        coef_type=np.array([141, 142, 132, 133, 121, 121, 121]),
        coef_order=np.array([0, 0, 0, 0, 1, 2, 3]),
        coef_val=np.array([io['C1'], io['C2'], io['B1'], io['B2'], io['A1'], io['A2'], io['A3']]),
        coef_fixed=np.array([io_fixed['C1'], io_fixed['C2'], io_fixed['B1'], io_fixed['B2'], io_fixed['A1'], io_fixed['A2'], io_fixed['A3']]),
        eo_val=np.array([images[i]['eo'] for i in img_order]),
        img_id=np.array(img_order),
        pt_ptr=np.array(ptr, np.int64),
        obj=np.array(obj, np.int32), xy=np.array(xy), sigma=np.array(sig),
        point_name=np.array(pt_order), point_xyz=np.array([points[n] for n in pt_order]),
        point_datum=np.array([len(n) <= 3 for n in pt_order]),
        bar_a=np.array([pidx[b[0]] for b in bars], np.int32), bar_b=np.array([pidx[b[1]] for b in bars], np.int32),
        bar_len=np.array([b[2] for b in bars]), bar_sigma=np.array([b[3] for b in bars]),
    )
    print('points', len(pt_order), 'images', len(img_order), 'image points', len(obj), 'bars', len(bars),
          'datum', int(sum(len(n) <= 3 for n in pt_order)), 'io', io, io_fixed)


if __name__ == '__main__':
    sys.exit(main())
