"""Out-of-bounds writes (-m gpu).  compute-sanitizer is not available on the GPU boxes of this pool, so the library has a
debugging mode of its own: with BSGPU_REDZONE=1 every device buffer it owns is allocated at exactly the size asked for
(normally 12.5 % head-room would absorb an overrun) and followed by a 4-KiB zone of a known pattern
(bs_call_b200/csrc/bsgpu_api.cu: DevBuf, bsgpu_debug_redzones).  The parity modules -- ragged sizes, empty inputs, one-site
windows, 700x panels, chunked pipelines, sessions -- are run once more like that in a child process (the switch is read when the
library is loaded); the zones are read back whenever a buffer is released and after the last test (tests/conftest.py)."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODULES = ["tests/test_gpu_parity.py", "tests/test_gpu_reader.py", "tests/test_gpu_writer.py", "tests/test_gpu_session.py",
           "tests/test_gpu_site_stats.py", "tests/test_gpu_profile.py", "tests/test_gpu_guard.py"]


def test_no_kernel_writes_behind_a_library_buffer():
    env = dict(os.environ)
    env["BSGPU_REDZONE"] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", *MODULES, "-m", "gpu", "-q", "-x", "-s", "-p", "no:cacheprovider"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=1500)
    tail = r.stdout[-3000:] + r.stderr[-1500:]
    m = re.search(r"redzones: checked (\d+) corrupt (\d+) rc (-?\d+)", r.stdout)
    assert m, tail
    print(m.group(0), "|", [ln for ln in r.stdout.split("\n") if " passed" in ln or " failed" in ln][-1:])
    assert r.returncode == 0, tail
    assert int(m.group(1)) > 50 and int(m.group(2)) == 0 and int(m.group(3)) == 1, tail
