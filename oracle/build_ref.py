"""Recipe: unpack the reference's own PoseLib build into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  The RePoseD hot path of kocurvik/mdrp lives in
PoseLib 2.0.5 + PR #152, which the reference ships as a prebuilt wheel
(/root/reference/demo/poselib-2.0.5-cp312-cp312-linux_x86_64.whl,
/root/reference/README.md:52).  There is no source to compile, so the
"reference build" is this wheel, unpacked (never copied into git history:
oracle/_ref/ is git-ignored, but it travels to the GPU box with the snapshot).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may load what this produces.
"""
import hashlib
import os
import sys
import zipfile

WHEEL = "/root/reference/demo/poselib-2.0.5-cp312-cp312-linux_x86_64.whl"
WHEEL_SHA256_PREFIX = "509a74fd"
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def have_ref() -> bool:
    return os.path.exists(os.path.join(DEST, "poselib", "__init__.py"))


def build(force: bool = False) -> bool:
    """Unpack the wheel if the reference tree is mounted.  Returns have_ref()."""
    if have_ref() and not force:
        return True
    if not os.path.exists(WHEEL):
        return have_ref()
    with open(WHEEL, "rb") as f:
        digest = hashlib.sha256(f.read()).hexdigest()
    if not digest.startswith(WHEEL_SHA256_PREFIX):
        raise RuntimeError(f"unexpected wheel sha256 {digest}")
    os.makedirs(DEST, exist_ok=True)
    zipfile.ZipFile(WHEEL).extractall(DEST)
    return have_ref()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "unavailable (no /root/reference)")
