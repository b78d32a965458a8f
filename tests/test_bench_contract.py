"""CPU: the parts of bench.py's JSON contract that can be checked without a GPU -- both arms describe the SAME workload (`config`), the
reference arm of the SDS configs reports `unavailable`, the profile-keyed traffic file parses and names the kernels the roofline block uses."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_workload_config_is_arm_independent_and_names_the_workload():
    import bench
    for cfg in ('cfg2', 'cfg3', 'cfg4', 'cfg5'):
        c = bench.workload_config(cfg, 4)
        assert c['workload'].startswith(cfg) and 'model' not in c and 'l2_flush' in c and c['parallelism'].endswith('dp4') or 'dp4' in c['parallelism']
    assert bench.workload_config('cfg2', 1) == bench.workload_config('cfg2', 1)
    assert len(bench.csrc_sha()) == 12


def test_reference_arm_of_sds_configs_reports_unavailable():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', 'cfg3'], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and 'unavailable' in line


def test_traffic_profile_names_the_roofline_kernels():
    p = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    assert os.path.exists(p)
    t = json.load(open(p))
    assert t['M'] == 4096 * 128 and len(t['csrc_sha']) == 12
    for k in ('fd_regulariser', 'field_fwd_main', 'field_bwd_sdf_tc_main', 'field_bwd_warp_tc'):
        assert t['kernels'][k] > 0
