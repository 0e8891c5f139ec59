"""Host logic of the streaming session (bs_call_b200/csrc/bsgpu_session.h) on the CPU: tests/session_harness.cpp drives it
with a stand-in for the device run of a batch -- random streams, batch sizes, slicings, one thread (non-blocking feeds,
reserve / commit) and two threads (feeder + printer, as in the reference; blocking feeds, and the bulk-input pattern: blocking
reserve, fill in place, commit, reserve again at once), cuts, rewind, truncated streams.  The results
of a session, concatenated, must be the stream; a deadlock ends the run through alarm()."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_session_host_logic(tmp_path):
    exe = str(tmp_path / "session_harness")
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "bs_call_b200", "csrc"), os.path.join(ROOT, "tests", "session_harness.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe, "150"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "session harness ok" in out.stdout
