#!/usr/bin/env python3
"""Run the UNMODIFIED reference orchestrator (tomography_3d_reconstruction.main(), /root/reference) on top of the drop-in
modules, on a GPU: the integration test of SURVEY.md section 7 step 1.

    python tools/run_orchestrator.py [--reference DIR] [--size 512] [--slices 20,64,20] [--log profiles/...log]

What this harness does -- and nothing else:
  * puts tomography_3d_reconstructor_b200/dropin ahead of the reference checkout on sys.path, so the orchestrator's
    `from voxel_processor import VoxelProcessor` etc. (tomography_3d_reconstruction.py:13-18) bind to the B200 classes;
    image_loader, visualizer, config and the orchestrator itself are the reference's own files, unmodified;
  * stubs the presentation packages that are absent from this image (matplotlib, plotly): recording stand-ins, so the
    calls the reference makes on them (go.Mesh3d(...), fig.write_html(...)) are seen and counted;
  * builds the BASELINE configs[0] fixture with the reference's OWN generator (cv2 ellipse base mask -> 64 copies in
    Section_1, simple_generator.generate_slices_from_mask end caps in Section_0 / Section_2);
  * points config.DATA_PATH / GLB_FILENAME / INTERACTIVE_HTML (private paths of the author's machine, config.py:22,46-47)
    at a temporary directory -- module attributes set from outside, the file is not touched;
  * runs main(), checks its return code, re-parses the GLB it wrote, and compares the printed volume with the CPU oracle.

--reference defaults to /root/reference; on a GPU box (where that path does not exist) pass a temporary copy.
"""
import argparse
import contextlib
import io
import os
import sys
import tempfile
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "tomography_3d_reconstructor_b200", "dropin")

CALLS = []


class _Recorder:
    """Stand-in for a presentation object: records how it is used, returns itself."""

    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        CALLS.append((self._name, sorted(k)))
        return self

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Recorder(self._name + "." + item)


def stub_module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    def _getattr(item, _n=name):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Recorder(_n + "." + item)
    m.__getattr__ = _getattr
    sys.modules[name] = m
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--slices", default="20,64,20")
    ap.add_argument("--log", default=None)
    args = ap.parse_args()
    s0, s1, s2 = (int(v) for v in args.slices.split(","))
    if not os.path.isfile(os.path.join(args.reference, "tomography_3d_reconstruction.py")):
        raise SystemExit("no reference checkout at %s" % args.reference)

    import cv2
    try:
        import matplotlib.pyplot  # noqa: F401
    except ImportError:
        stub_module("matplotlib")
        stub_module("matplotlib.pyplot")
    try:
        import plotly.graph_objects  # noqa: F401
    except ImportError:
        stub_module("plotly")
        stub_module("plotly.graph_objects")
    sys.path.insert(0, args.reference)
    sys.path.insert(0, DROPIN)
    sys.path.insert(0, ROOT)

    log = io.StringIO()

    class Tee:
        def write(self, s):
            log.write(s)
            sys.__stdout__.write(s)

        def flush(self):
            sys.__stdout__.flush()

    with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(Tee()):
        n = args.size
        base = np.zeros((n, n), dtype=np.uint8)
        cv2.ellipse(base, (n // 2, n // 2), (int(n * 200 / 512), int(n * 140 / 512)), 0, 0, 360, 255, -1)
        sec1 = os.path.join(d, "Section_1")
        os.makedirs(sec1)
        for k in range(1, s1 + 1):
            cv2.imwrite(os.path.join(sec1, "Mask_Patient_%d.png" % k), base)
        import simple_generator                              # the reference's generator
        with contextlib.redirect_stdout(io.StringIO()):
            simple_generator.generate_slices_from_mask(os.path.join(sec1, "Mask_Patient_1.png"), s0, os.path.join(d, "Section_0"), 1, False)
            simple_generator.generate_slices_from_mask(os.path.join(sec1, "Mask_Patient_%d.png" % s1), s2, os.path.join(d, "Section_2"),
                                                       s1, True)
        import config                                        # the reference's config, paths redirected
        config.DATA_PATH = d
        config.GLB_FILENAME = os.path.join(d, "tomography_model.glb")
        config.INTERACTIVE_HTML = os.path.join(d, "tomography_3d_interactive.html")
        import tomography_3d_reconstruction as orch          # the reference's orchestrator
        import voxel_processor
        import surface_extractor
        import volume_calculator
        import glb_exporter
        bound = {m.__name__: os.path.relpath(m.__file__, ROOT) if m.__file__.startswith(ROOT) else m.__file__
                 for m in (orch, voxel_processor, surface_extractor, volume_calculator, glb_exporter, sys.modules["image_loader"],
                           sys.modules["visualizer"], config)}
        print("[harness] modules:", bound)
        assert bound["voxel_processor"].startswith("tomography_3d_reconstructor_b200")
        assert os.path.abspath(orch.__file__).startswith(os.path.abspath(args.reference))
        import torch
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = orch.main()
        torch.cuda.synchronize()
        t_main = time.perf_counter() - t0
        print("[harness] main() returned %r in %.3f s (first call: includes CUDA context / library load)" % (rc, t_main))
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            rc2 = orch.main()
        torch.cuda.synchronize()
        t_again = time.perf_counter() - t0
        print("[harness] second main() returned %r in %.3f s" % (rc2, t_again))
        assert rc == 0 and rc2 == 0
        # what the presentation layer was asked to do
        names = [c[0] for c in CALLS]
        print("[harness] presentation calls:", sorted(set(names)))
        # the GLB the orchestrator wrote through the drop-in exporter
        from tomography_3d_reconstructor_b200.glb_exporter import parse_glb
        pos, idx, col, doc = parse_glb(open(config.GLB_FILENAME, "rb").read())
        red, blue = int((col[:, 0] == 255).sum()), int((col[:, 2] == 255).sum())
        print("[harness] GLB: %d vertices, %d faces, %d red / %d blue highlighted vertices, %d bytes" %
              (len(pos), len(idx), red, blue, os.path.getsize(config.GLB_FILENAME)))
        assert len(pos) > 1000 and len(idx) > 1000 and red > 0 and blue > 0
        # the same stack through the CPU oracle (the reference's arithmetic with skimage present)
        from oracle import cpu_ref
        loader = sys.modules["image_loader"].ImageLoader()
        with contextlib.redirect_stdout(io.StringIO()):
            loader.load_mask_images(d, config.THRESHOLD, config.LOAD_SIDES)
        masks = np.stack(loader.mask_images).astype(np.uint8) * 255
        t0 = time.perf_counter()
        ref = cpu_ref.reference_pipeline(masks, 200, (s0, s1, s2), config.TOTAL_DEPTH_MM, config.X_LENGTH_MM, config.Y_LENGTH_MM)
        t_oracle = time.perf_counter() - t0
        line = [ln for ln in log.getvalue().splitlines() if ln.startswith("Volume:")][0]
        vol_printed = float(line.split()[1])
        print("[harness] oracle: one pass %.2f s (the reference's main() repeats smoothing 5x and extraction 4x); mesh volume %.4f mm3, "
              "orchestrator printed %.4f mm3" % (t_oracle, ref["mesh_volume"], vol_printed))
        assert abs(vol_printed - ref["mesh_volume"]) <= 1e-3 + 1e-5 * ref["mesh_volume"]
        assert np.array_equal(pos.view(np.uint32), ref["vertices"].view(np.uint32)) and len(idx) == len(ref["faces"])
        print("[harness] OK: GLB vertices bit-equal to the oracle's, face count equal, volume equal")
    if args.log:
        os.makedirs(os.path.dirname(os.path.abspath(args.log)), exist_ok=True)
        open(args.log, "w").write(log.getvalue())


if __name__ == "__main__":
    main()
