"""Drives the Redis module (redis_hnsw_b200/libredis_hnsw_b200.so) inside tests/fake_redis/fake_redis_host: one
process = one "server boot"; commands go in as a script, replies come back as one JSON value per command."""
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tests", "fake_redis", "fake_redis_host")
MODULE = os.path.join(ROOT, "redis_hnsw_b200", "libredis_hnsw_b200.so")


def build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "redis_hnsw_b200", "csrc", "redis")], stdout=subprocess.DEVNULL)


def fmt(v):
    """A float the way redis-cli would carry it: shortest repr that round-trips the f32."""
    return repr(float(v))


def run(commands, timeout=600, env=None):
    """commands: list of strings or lists of words.  Returns the list of decoded replies (same length).
    `env`: extra environment for the host process (e.g. FAKE_REDIS_NO_EVENTS=1)."""
    build()
    lines = [c if isinstance(c, str) else " ".join(str(w) for w in c) for c in commands]
    p = subprocess.run([HOST, MODULE], input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=timeout,
                       env=dict(os.environ, **(env or {})))
    assert p.returncode == 0, "fake_redis_host rc=%d stderr=%s" % (p.returncode, p.stderr[-2000:])
    out = [json.loads(l) for l in p.stdout.splitlines() if l.strip()]
    assert len(out) == len(lines), "expected %d replies, got %d\n%s" % (len(lines), len(out), p.stderr[-2000:])
    return out


def is_error(reply):
    return isinstance(reply, dict) and "error" in reply


def pairs(reply):
    """Flat [k, v, k, v, ...] reply -> dict."""
    return {reply[i]: reply[i + 1] for i in range(0, len(reply), 2)}
