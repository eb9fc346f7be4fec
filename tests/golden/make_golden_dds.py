"""Writes tests/golden/dds_headers.json: the DDS file headers the REFERENCE's own SaveDDSTextureToFile lines
(ScreenGrab11.cpp:72-208,819-906, compiled by oracle/build_ref.sh into oracle/_ref) produce for the formats the mod can
capture.  Run where /root/reference exists:  python tests/golden/make_golden_dds.py"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import pyoracle as po  # noqa: E402

BPT = {po.FMT_RGBA8: 4, po.FMT_BGRA8: 4, po.FMT_RGBA16F: 8, po.FMT_RGBA32F: 16, po.FMT_RGB10A2: 4}
cases = []
for fmt, bpt in BPT.items():
    for (w, h) in ((3, 2), (37, 21), (2244, 2492)):
        cases.append({"format": fmt, "width": w, "height": h, "bytes_per_texel": bpt, "header_hex": po.ref_dds_header(w, h, fmt, bpt).hex()})
(Path(__file__).parent / "dds_headers.json").write_text(json.dumps(cases, indent=1))
print(f"wrote {len(cases)} headers")
