"""Shared helpers for the test-suite (seeded sequence generators, golden loading)."""
import argparse
import json
import os
import random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def ns(**kw):
    return argparse.Namespace(**kw)


def expected_of(case):
    return [tuple(r) for r in case["result"]]


_EXC = {"ValueError": ValueError, "IndexError": IndexError, "AssertionError": AssertionError,
        "AttributeError": AttributeError}


def exc_of(case):
    return _EXC[case["raises"]]


ALPHABETS = ["ACGT", "ACGT", "ACGTN", "AC", "ACGTacgtNn", "AT"]


def random_seq(rng, n, alphabets=ALPHABETS, exotic=False):
    """Repeat-rich random sequence of exactly n symbols."""
    alpha = rng.choice(alphabets)
    parts, total = [], 0
    while total < n:
        kind = rng.random()
        if kind < 0.45:
            piece = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 40)))
        elif kind < 0.88:
            unit = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 60 if rng.random() < 0.2 else 9)))
            piece = unit * rng.randint(1, 14) + unit[:rng.randint(0, len(unit))]
        elif kind < 0.97 or not exotic:
            piece = rng.choice("Nn") * rng.randint(1, 70)
        else:
            piece = rng.choice("RYKMSWBDHV") * rng.randint(1, 12)
        parts.append(piece)
        total += len(piece)
    return "".join(parts)[:n]
