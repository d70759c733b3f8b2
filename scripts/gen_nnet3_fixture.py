"""Extract the nnet3 text fixture the reference's own parser test uses (internal/nnet/weight_loader_test.go:11-52, `testComponents`:
"Test data from actual nnet3-copy --binary=false output") into tests/golden/nnet3_text_fixture.txt.  Run in the build container
(where /root/reference is mounted); the fixture travels, the reference does not."""
import re
from pathlib import Path

src = Path("/root/reference/internal/nnet/weight_loader_test.go").read_text()
m = re.search(r"const testComponents = `(.*?)`", src, re.S)
assert m, "fixture not found"
out = Path(__file__).resolve().parent.parent / "tests" / "golden" / "nnet3_text_fixture.txt"
out.write_text(m.group(1))
print(f"wrote {out} ({len(m.group(1))} bytes)")
