"""Synthetic inputs for the treatment-planning front end (configs C3-C5 of BASELINE.json): a head-like
CT volume as .mha, a generic PBS beam-model file in the reference's `pbs:<file>` format
(moqui/base/mqi_treatment_machine_pbs.hpp:400-516), a text plan (csrc/mqi_tps_host.hpp) and the moqui
input-parameter file.  Everything is seeded; there is no network for real CT / plan data.
"""
import os

import numpy as np


def head_ct(n=(512, 512, 200), spacing=(1.0, 1.0, 2.5), seed=1):
    """HU volume [nz, ny, nx] int16: ellipsoid of soft tissue (0..60 HU) with a skull shell
    (700..1200 HU), two air cavities (-1000) and air outside."""
    nx, ny, nz = n
    rng = np.random.default_rng(seed)
    x = (np.arange(nx) - (nx - 1) / 2.0) * spacing[0]
    y = (np.arange(ny) - (ny - 1) / 2.0) * spacing[1]
    z = (np.arange(nz) - (nz - 1) / 2.0) * spacing[2]
    ext = np.array([nx * spacing[0], ny * spacing[1], nz * spacing[2]]) / 2.0
    a, b, c = 0.36 * ext[0], 0.44 * ext[1], 0.48 * ext[2]
    hu = np.full((nz, ny, nx), -1000, dtype=np.int16)
    yy, xx = np.meshgrid(y, x, indexing="ij")
    for k in range(nz):
        r2 = (xx / a) ** 2 + (yy / b) ** 2 + (z[k] / c) ** 2
        inner = r2 < 0.90 ** 2
        shell = (r2 < 1.0) & ~inner
        sl = hu[k]
        sl[inner] = rng.integers(0, 61, size=int(inner.sum()), dtype=np.int16)
        sl[shell] = rng.integers(700, 1201, size=int(shell.sum()), dtype=np.int16)
        for (cx, cy, cz, r) in ((0.25 * a, -0.45 * b, 0.1 * c, 0.12 * a), (-0.25 * a, -0.45 * b, 0.1 * c, 0.12 * a)):
            cav = (xx - cx) ** 2 + (yy - cy) ** 2 + (z[k] - cz) ** 2 < r * r
            sl[cav & inner] = -1000
    origin = (x[0], y[0], z[0])
    return hu, origin


def write_mha(path, hu, origin, spacing, element="MET_FLOAT"):
    """MetaImage with float voxels, the layout tps_env::read_ct_image parses (mqi_tps_env.hpp:615-701); MET_SHORT is
    the other element type the B200 front end accepts."""
    nz, ny, nx = hu.shape
    hdr = ("ObjectType = Image\nNDims = 3\nBinaryData = True\nBinaryDataByteOrderMSB = False\nCompressedData = False\n"
           "TransformMatrix = 1 0 0 0 1 0 0 0 1\nOffset = %.9g %.9g %.9g\nCenterOfRotation = 0 0 0\n"
           "AnatomicalOrientation = RAI\nElementSpacing = %.9g %.9g %.9g\nDimSize = %d %d %d\n"
           "ElementType = %s\nElementDataFile = LOCAL\n"
           % (origin[0], origin[1], origin[2], spacing[0], spacing[1], spacing[2], nx, ny, nz, element))
    with open(path, "wb") as f:
        f.write(hdr.encode())
        f.write(hu.astype({"MET_FLOAT": np.float32, "MET_SHORT": np.int16, "MET_UCHAR": np.uint8}[element]).tobytes())


def write_mask_mha(path, mask, origin=(0.0, 0.0, 0.0), spacing=(1.0, 1.0, 1.0)):
    """uint8 0/1 mask volume [nz, ny, nx] in the layout mask_reader::read_mha_file parses
    (moqui/base/mqi_file_handler.hpp:38-99)."""
    nz, ny, nx = mask.shape
    hdr = ("ObjectType = Image\nNDims = 3\nBinaryData = True\nBinaryDataByteOrderMSB = False\nCompressedData = False\n"
           "TransformMatrix = 1 0 0 0 1 0 0 0 1\nOffset = %.9g %.9g %.9g\nCenterOfRotation = 0 0 0\n"
           "AnatomicalOrientation = RAI\nElementSpacing = %.9g %.9g %.9g\nDimSize = %d %d %d\n"
           "ElementType = MET_UCHAR\nElementDataFile = LOCAL\n"
           % (origin[0], origin[1], origin[2], spacing[0], spacing[1], spacing[2], nx, ny, nz))
    with open(path, "wb") as f:
        f.write(hdr.encode())
        f.write(np.ascontiguousarray(mask, dtype=np.uint8).tobytes())


def write_structures(path, rois):
    """Text structure set (csrc/mqi_tps_host.hpp text_structures): rois = {name: [contour, ...]}, a contour is
    a list of (x, y, z) points of one closed planar polygon (the RTSTRUCT's ContourData)."""
    with open(path, "w") as f:
        for name, contours in rois.items():
            f.write("[roi]\nname %s\n" % name)
            for c in contours:
                f.write("[contour]\n# x y z\n")
                for (x, y, z) in c:
                    f.write("%.7g %.7g %.7g\n" % (x, y, z))


def ellipse_contours(a, b, z_values, n_points=48, centre=(0.0, 0.0), z_scale=None):
    """One elliptical contour per z (semi-axes shrink with z_scale(z) if given)."""
    out = []
    for z in z_values:
        k = 1.0 if z_scale is None else z_scale(z)
        t = np.linspace(0.0, 2.0 * np.pi, n_points, endpoint=False)
        out.append([(centre[0] + k * a * np.cos(u), centre[1] + k * b * np.sin(u), float(z)) for u in t])
    return out


def beam_model_rows(e_lo=60.0, e_hi=240.0, step=10.0):
    """[spot] rows: Enominal E dE x y xp yp ratio"""
    rows = []
    e = e_lo
    while e <= e_hi + 1e-6:
        t = (e - e_lo) / (e_hi - e_lo)
        rows.append((e, e + 0.2, 0.006 * e, 6.0 - 3.0 * t, 6.5 - 3.0 * t, 0.004 - 0.002 * t, 0.0042 - 0.002 * t,
                     1.0e6 * (1.0 + 0.5 * t)))
        e += step
    return rows


def write_beam_model(path, sad=(2000.0, 1800.0)):
    with open(path, "w") as f:
        f.write("# generic PBS beam model (synthetic)\n[geometry]\n")
        f.write("SAD(mm) %g %g\nrangeshifter(mm) 300 300\nrangeshifter_snout_gap(mm) 10\naperture(mm) 300 300\n" % sad)
        f.write("[rangeshifter_thickness]\n\"RS1\" 40.0\n")
        f.write("[spot]\n# Enominal E dE x y xp yp ratio\n")
        for r in beam_model_rows():
            f.write("%g %g %g %g %g %g %g %g\n" % r)
        f.write("[time]\nE(sec) 1.0\nI(nA) 2.0\ndT_up(ms) 0.1\ndT_down(ms) 0.1\nset_x(ms) 1\nset_y(ms) 1\n"
                "velocity_x(m/sec) 10\nvelocity_y(m/sec) 10\n")


def spot_list(n_layers=25, e_lo=70.0, e_hi=180.0, pitch=5.0, half_width=22.0, seed=1, meterset_scale=1.0):
    """~n_layers energy layers of spots on a `pitch` mm grid inside a disc, log-normal metersets."""
    rng = np.random.default_rng(seed)
    spots = []
    g = np.arange(-half_width, half_width + 1e-6, pitch)
    for e in np.linspace(e_lo, e_hi, n_layers):
        for x in g:
            for y in g:
                if x * x + y * y <= half_width * half_width + 1e-6:
                    spots.append((float(e), float(x), float(y), float(meterset_scale * rng.lognormal(0.0, 0.5))))
    return spots


def write_plan(path, beams, name="synthetic", fractions=30):
    """beams: list of dicts {name, gantry, couch, collimator, iso, snout, spots}"""
    with open(path, "w") as f:
        f.write("[plan]\nname %s\nfractions %d\n" % (name, fractions))
        for b in beams:
            f.write("[beam]\nname %s\ngantry_angle %g\ncouch_angle %g\ncollimator_angle %g\n" %
                    (b["name"], b.get("gantry", 0.0), b.get("couch", 0.0), b.get("collimator", 0.0)))
            iso = b.get("iso", (0.0, 0.0, 0.0))
            f.write("isocenter %g %g %g\nsnout_position %g\n" % (iso[0], iso[1], iso[2], b.get("snout", 250.0)))
            # optional beamline: range shifter by ID (thickness from the beam model) or by WET, aperture blocks
            if b.get("rangeshifter_ids"):
                f.write("rangeshifter_id %s\n" % " ".join(b["rangeshifter_ids"]))
            if b.get("rangeshifter_wet"):
                f.write("rangeshifter_wet %g %g\n" % tuple(b["rangeshifter_wet"]))
            if b.get("blocks"):
                f.write("block_thickness %g\nblock_tray_distance %g\n" % (b.get("block_thickness", 20.0), b.get("block_tray_distance", 60.0)))
                for poly in b["blocks"]:
                    f.write("[block]\n# x y\n")
                    for (x, y) in poly:
                        f.write("%.6g %.6g\n" % (x, y))
            f.write("[spots]\n# E x y meterset\n")
            for s in b["spots"]:
                f.write("%.6g %.6g %.6g %.8g\n" % s)


def write_input(path, parent_dir, out_dir, **kw):
    keys = {
        "GPUID": "0", "RandomSeed": "12345", "UseAbsolutePath": "false", "TotalThreads": "-1", "MaxHistoriesPerBatch": "0",
        "ParentDir": parent_dir, "DicomDir": ".", "CTVolumeName": "ct.mha", "PlanFile": "plan.txt",
        "Scorer": "Dose", "SourceType": "FluenceMap", "SimulationType": "perBeam", "BeamNumbers": "0",
        "ParticlesPerHistory": "1000000", "ScoreToCTGrid": "true", "OutputDir": out_dir, "OutputFormat": "raw",
        "OverwriteResults": "true", "RBE": "1.1", "NumberOfFraction": "30", "Machine": "pbs:machine.txt", "Calibration": "default",
    }
    keys.update({k: str(v) for k, v in kw.items()})
    with open(path, "w") as f:
        f.write("## synthetic moqui input\n")
        for k, v in keys.items():
            f.write("%s %s\n" % (k, v))


def make_case(root, n=(64, 64, 40), spacing=(4.0, 4.0, 6.0), n_layers=4, pitch=10.0, half_width=20.0, beams=1, seed=1,
              beam_extra=None, **input_kw):
    """Write ct.mha, machine.txt, plan.txt and moqui_tps.in under `root`; returns the input-file path."""
    os.makedirs(root, exist_ok=True)
    hu, origin = head_ct(n, spacing, seed)
    write_mha(os.path.join(root, "ct.mha"), hu, origin, spacing)
    write_beam_model(os.path.join(root, "machine.txt"))
    bl = []
    for i in range(beams):
        bl.append({"name": "G%03d" % (90 * i), "gantry": 90.0 * i, "couch": 0.0, "collimator": 0.0, "iso": (0.0, 0.0, 0.0),
                   "snout": 250.0, "spots": spot_list(n_layers=n_layers, pitch=pitch, half_width=half_width, seed=seed + i)})
        bl[-1].update(beam_extra or {})
    write_plan(os.path.join(root, "plan.txt"), bl)
    out = input_kw.pop("OutputDir", os.path.join(root, "out"))
    inp = os.path.join(root, "moqui_tps.in")
    write_input(inp, root, out, **input_kw)
    return inp
