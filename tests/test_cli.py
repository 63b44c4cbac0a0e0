"""The `chromobius_b200` command line tool (chromobius_b200/csrc/main_cli.cc) against the reference CLI's own tests
(commands/main_predict.test.cc:22-44, commands/main_benchmark.test.cc:22-58) and against the Python API for the b8 / 01
record formats. Argument handling runs without a GPU; decoding needs one (there is no CPU fallback)."""
import os
import subprocess

import numpy as np
import pytest

from dem_util import read_dem, NumpyDemSampler

TINY_DEM = """
        error(0.1) D0 L0
        error(0.1) D0 D1 L1
        error(0.1) D1 L2
        detector(0, 0, 0, 0) D0
        detector(0, 0, 0, 1) D1
"""


@pytest.fixture(scope='module')
def cli(built):
    import __graft_entry__ as g
    return g.build_cli()


def run(cli, args, stdin=b''):
    p = subprocess.run([cli] + list(args), input=stdin, capture_output=True, timeout=300)
    return p.returncode, p.stdout, p.stderr.decode()


def test_help_and_argument_errors(cli):
    rc, out, _ = run(cli, ['help'])
    assert rc == 0 and b'chromobius_b200 predict' in out and b'chromobius_b200 benchmark' in out
    rc, _, err = run(cli, [])
    assert rc == 1 and err.startswith('Specify a chromobius_b200 command to run.')
    rc, _, err = run(cli, ['frobnicate'])
    assert rc == 1 and "Unrecognized chromobius_b200 command 'frobnicate'." in err
    rc, _, err = run(cli, ['predict', '--nope'])
    assert rc == 1 and 'Unrecognized command line argument --nope for `chromobius_b200 predict`' in err
    rc, _, err = run(cli, ['predict', '--in_format', 'xyz', '--dem', '/dev/null'])
    assert rc == 1 and "Unrecognized value 'xyz' for enum flag '--in_format'" in err
    rc, _, err = run(cli, ['predict'])
    assert rc == 1 and "--dem" in err


def test_benchmark_needs_observables(cli, tmp_path):
    dem = tmp_path / 'd.dem'
    dem.write_text(TINY_DEM)
    rc, _, err = run(cli, ['benchmark', '--dem', str(dem)])
    assert rc == 1 and 'Must specify --in_includes_appended_observables or --obs_in.' in err


def test_compile_errors_reach_stderr(cli, tmp_path):
    dem = tmp_path / 'bad.dem'
    dem.write_text('error(0.1) D0 L0\n')
    rc, _, err = run(cli, ['predict', '--dem', str(dem)])
    assert rc == 1 and 'Expected all detectors to have at least 4 coordinates' in err


@pytest.mark.gpu
def test_predict_reference_known_answer(cli, tmp_path):
    dem = tmp_path / 'd.dem'
    dem.write_text(TINY_DEM)
    rc, out, err = run(cli, ['predict', '--dem', str(dem), '--in_format', 'dets', '--out_format', 'dets'],
                       b'shot\nshot D0\nshot D1\nshot D0 D1')
    assert rc == 0, err
    assert out == b'shot\nshot L0\nshot L2\nshot L1\n'
    # the same shots through 01 in / b8 out and b8 in / hits out
    rc, out, err = run(cli, ['predict', '--dem', str(dem), '--in_format', '01', '--out_format', 'b8'], b'00\n10\n01\n11\n')
    assert rc == 0 and out == bytes([0, 1, 4, 2]), err
    rc, out, err = run(cli, ['predict', '--dem', str(dem), '--in_format', 'b8', '--out_format', 'hits'], bytes([0, 1, 2, 3]))
    assert rc == 0 and out == b'\n0\n2\n1\n', err
    # appended observables are ignored by predict
    rc, out, err = run(cli, ['predict', '--dem', str(dem), '--in_format', '01', '--in_includes_appended_observables'],
                       b'00111\n10000\n01010\n11101\n')
    assert rc == 0 and out == b'000\n100\n001\n010\n', err


@pytest.mark.gpu
def test_benchmark_reference_known_answer(cli, tmp_path):
    dem = tmp_path / 'd.dem'
    dem.write_text(TINY_DEM)
    rc, out, err = run(cli, ['benchmark', '--dem', str(dem), '--in_format', 'dets', '--in_includes_appended_observables'],
                       b'shot L0\nshot D0 L0\nshot D1 L2\nshot D0 D1 L1')
    assert rc == 0, err
    expected_prefix = """                                num_shots = 4
                             num_mistakes = 1
                        mistakes_per_shot = 0.25

                     num_detection_events = 4
                   num_detectors_per_shot = 2
                       detection_fraction = 0.5

                            setup_seconds ="""
    assert out.decode().startswith(expected_prefix)
    # observables from a separate file
    obs = tmp_path / 'obs.01'
    obs.write_text('100\n100\n001\n010\n')
    rc, out2, err = run(cli, ['benchmark', '--dem', str(dem), '--in_format', '01', '--obs_in', str(obs)], b'00\n10\n01\n11\n')
    assert rc == 0 and 'num_mistakes = 1' in out2.decode(), err


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['midout_color_code_d9_r36_p1000', 'phenom_color_code_d5_r5_p1000_with_ignored'])
def test_predict_b8_matches_python_api(cli, tmp_path, name):
    import chromobius_b200 as cb
    text = read_dem(name)
    dem = tmp_path / 'c.dem'
    dem.write_text(text)
    dets, obs = NumpyDemSampler(text).sample(5000, seed=3)
    dec = cb._compile_text(text, device=0)
    want = dec.predict_obs_flips_from_dets_bit_packed(dets)
    b8_in = tmp_path / 'dets.b8'
    dets.tofile(b8_in)
    out_path = tmp_path / 'pred.b8'
    rc, _, err = run(cli, ['predict', '--dem', str(dem), '--in', str(b8_in), '--in_format', 'b8', '--out', str(out_path),
                           '--out_format', 'b8', '--batch', '1024'])
    assert rc == 0, err
    got = np.fromfile(out_path, dtype=np.uint8).reshape(want.shape)
    assert (got == want).all()
    rc, out, err = run(cli, ['predict', '--dem', str(dem), '--in', str(b8_in), '--in_format', 'b8', '--out_format', '01'])
    assert rc == 0, err
    lines = out.decode().split('\n')[:-1]
    assert len(lines) == len(dets)
    nobs = dec.num_observables
    for k in (0, 1, 17, len(dets) - 1):
        assert lines[k] == ''.join('1' if (want[k, b >> 3] >> (b & 7)) & 1 else '0' for b in range(nobs))
    # describe_decoder prints the same dump as the Python API
    rc, out, err = run(cli, ['describe_decoder', '--in', str(dem)])
    assert rc == 0 and out.decode().rstrip('\n') == dec.describe().rstrip('\n'), err
