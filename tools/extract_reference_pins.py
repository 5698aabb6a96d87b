"""Regenerates tests/golden/tprob.json: the 216 bytes of the Mendelian transmission table _TPROB
embedded in the reference binary (Mach-O __TEXT,__const, file offset 0x35b10 = VM 0x100035b10).
This is the only numerical pin the reference offers. Needs /root/reference, so it
runs only in the build container; the committed JSON is what the tests read."""
import json
from pathlib import Path

BIN = Path("/root/reference/exec/pedigraph")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "tprob.json"

if __name__ == "__main__":
    d = BIN.read_bytes()
    assert d[:4] == bytes.fromhex("cffaedfe"), "not the Mach-O x86_64 binary the survey describes"
    raw = d[0x35B10:0x35B10 + 216]
    OUT.write_text(json.dumps({
        "source": "ngthomas/pedFac exec/pedigraph (Mach-O x86_64), __TEXT,__const, file offset 0x35b10, "
                  "symbol _TPROB, 27 little-endian doubles indexed gk*9+gm*3+gp",
        "generated_by": "tools/extract_reference_pins.py",
        "hex": raw.hex()}, indent=1))
    print("wrote", OUT)
