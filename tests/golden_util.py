"""Loads tests/golden/proofs.json (made by tools/make_golden.py) and rebuilds the circuits it names."""
import json
import os

from tests.util import pkg

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proofs.json")


def load():
    return json.load(open(_PATH))["cases"]


def circuit_of(entry):
    return getattr(pkg().synth, entry["generator"])(*entry["args"], **entry["kwargs"])


def points_of(lst):
    return [None if p is None else (int(p[0], 16), int(p[1], 16)) for p in lst]
