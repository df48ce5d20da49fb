"""Writes tests/golden/reference_pins.json: sha256 of every reference line range the oracle restates
(tests/test_reference_pin.py::RANGES). Run in a container that mounts the reference at /root/reference:
    python tests/golden/make_reference_pins.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_reference_pin as t  # noqa: E402

pins = {"reference": "gong9/rag-era mounted at " + t.REF,
        "sha256": {f"{rel}:{a}-{b}": t.sha(rel, a, b) for rel, a, b in t.RANGES}}
json.dump(pins, open(t.PINS, "w"), indent=1)
print("wrote", t.PINS)
