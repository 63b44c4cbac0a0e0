"""CPU tier: the warp-synchronous blossom solver (Blossom::solve_w, opt-in with -DCB_WARP_SOLVER) against the sequential
one (Blossom::solve, what the kernels run by default and what the oracle restates) on random components. 32 host threads
stand in for the lanes of a warp; matching, labels and pair counts must be identical (tests/sim/warp_blossom_check.cc)."""
import os
import subprocess

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')


def test_warp_solver_takes_the_same_steps_as_the_sequential_solver(tmp_path):
    exe = str(tmp_path / 'warp_blossom_check')
    subprocess.check_call(['g++', '-O1', '-std=c++17', '-pthread', '-Wno-sign-compare', '-o', exe,
                           os.path.join(ROOT, 'tests', 'sim', 'warp_blossom_check.cc')])
    out = subprocess.run([exe, '400', '7'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    assert ' 0 failures' in out.stdout
    # the sample must exercise blossoms and components without a perfect matching
    words = out.stdout.strip().split('\n')[-1].replace('(', ' ').split()
    assert int(words[2]) > 100 and int(words[5]) > 50, out.stdout
