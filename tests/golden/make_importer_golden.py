"""Writes tests/golden/importers.json from the compiled reference (oracle/_ref): python tests/golden/make_importer_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]
import refbind  # noqa: E402
from importer_cases import CASES  # noqa: E402
from test_importers import digest, reference  # noqa: E402

ref = refbind.Ref("scalar")
out = {}
for name, kind, text, options in CASES:
    pts, polys, parts, filter_, names = reference(ref, kind, text, options)
    if kind == "ply":
        parts = [("Imported", len(polys))]  # dfpsr_import_ply's single part
    out[name] = {"counts": [len(pts), len(polys)], "texture_names": names, "sha256": digest(pts, polys, parts, filter_)}
json.dump(out, open(os.path.join(HERE, "importers.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
