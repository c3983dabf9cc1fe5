"""Extract the golden vectors the reference's own tests hold for the hot path into JSON fixtures.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Sources (relative to the reference checkout):
  G1  test/test_spline_grid.jl:29-53   5x6x2 control points -> 7x9x2 evaluated grid (Float32 grid)
  G2  test/test_local_refinement.jl:62-66   18x2 Int32 refinement index set
The literals are parsed from the Julia source text; nothing is executed.
"""
import json
import re
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def parse_julia_array(text: str):
    """Parse a Julia ``[a b; c d;;; e f; g h]`` literal into nested lists [plane][row][col]."""
    text = text.strip()
    assert text.startswith("[") and text.endswith("]")
    planes = []
    for plane in text[1:-1].split(";;;"):
        rows = []
        for row in plane.split(";"):
            vals = [float(v) for v in row.split()]
            if vals:
                rows.append(vals)
        planes.append(rows)
    return planes


def bracket_literals(src: str):
    """All top-level ``[...]`` literals containing ';;;' (the 3-d arrays)."""
    out, depth, start = [], 0, None
    for i, ch in enumerate(src):
        if ch == "[":
            if depth == 0:
                start = i
            depth += 1
        elif ch == "]":
            depth -= 1
            if depth == 0 and ";;;" in src[start:i + 1]:
                out.append(src[start:i + 1])
    return out


def main():
    src = (REF / "test/test_spline_grid.jl").read_text()
    lits = bracket_literals(src)
    assert len(lits) == 2, len(lits)
    cp = parse_julia_array(lits[0])       # [2][5][6]
    ev = parse_julia_array(lits[1])       # [2][7][9]
    assert (len(cp), len(cp[0]), len(cp[0][0])) == (2, 5, 6)
    assert (len(ev), len(ev[0]), len(ev[0][0])) == (2, 7, 9)
    g1 = {
        "source": "test/test_spline_grid.jl:29-53",
        "n_control_points": [5, 6], "degree": [3, 2], "n_sample_points": [7, 9], "Nout": 2,
        "float_type": "Float32",
        "layout": "[o][i1][i2] nested lists: value at Julia index (i1+1, i2+1, o+1)",
        "control_points": cp, "eval": ev,
    }
    (OUT / "g1_spline_grid_5x6.json").write_text(json.dumps(g1, indent=1))

    src = (REF / "test/test_local_refinement.jl").read_text()
    m = re.search(r"Int32\[(.*?)\]", src, re.S)
    rows = [[int(v) for v in r.split()] for r in m.group(1).replace("\n", " ").split(";")]
    assert len(rows) == 18 and all(len(r) == 2 for r in rows)
    g2 = {
        "source": "test/test_local_refinement.jl:48-67",
        "n_control_points": [6, 6], "degree": [2, 2], "n_sample_points": [50, 50], "Nout": 3,
        "float_type": "Float32",
        "error_block": {"dim1": [20, 40], "dim2": [10, 30], "out": 2, "note": "1-based inclusive"},
        "refinement_indices": rows,
    }
    (OUT / "g2_error_informed_indices.json").write_text(json.dumps(g2, indent=1))
    print("wrote", [p.name for p in OUT.glob("*.json")])


if __name__ == "__main__":
    main()
