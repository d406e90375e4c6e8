"""bench.py's CPU arm (baseline/reference_arm.py): runs the UNMODIFIED reference when it is available (the build
container's /root/reference, or the baseline/_ref copy on the GPU box), the oracle port otherwise. CPU-only tests."""
import os

import numpy as np
import pytest

from baseline import install_reference, ref_loader, reference_arm


def test_manifest_matches_the_copy():
    if not os.path.isdir(os.path.join(install_reference.DEST, "audiblelight")):
        pytest.skip("baseline/_ref not installed (run __graft_entry__.build() where /root/reference exists)")
    assert install_reference.verify()
    # and the copy is byte-identical to the source tree when that is around
    src = "/root/reference/audiblelight/synthesize.py"
    if os.path.exists(src):
        assert open(src, "rb").read() == open(os.path.join(install_reference.DEST, "audiblelight", "synthesize.py"), "rb").read()


def test_reference_arm_runs_the_unmodified_functions_on_a_static_scene():
    if not reference_arm.reference_installed():
        pytest.skip("reference not available")
    out = reference_arm.run(n_workers=1, workload="c1")
    assert out["kind"] == "reference" and out["cores"] == 1 and out["value"] > 0
    assert "no scaling" in out["sample"] or "in full" in out["sample"]
    syn = ref_loader.load_reference_synthesize()
    assert getattr(syn, "_alr_is_reference", False)
    assert os.path.abspath(syn.__file__).startswith(os.path.abspath(ref_loader.REFERENCE_ROOT))


def test_reference_arm_never_loads_the_cuda_library():
    """The arm imports audiblelight_b200.workload for the scene specs only: the C-ABI library must stay unloaded."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from baseline import reference_arm\n"
            "import audiblelight_b200.workload\n"
            "from audiblelight_b200 import _lib\n"
            "assert _lib._lib is None, 'libalrender.so was loaded'\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libalrender' not in maps\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, "-c", code], check=True)
