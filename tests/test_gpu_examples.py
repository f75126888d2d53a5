"""The example scripts (the reference's examples/feedforward_mnist.py, char_rnn.py and
higher_order_deriv_visualization.py rewritten on synthetic data) run end to end on the GPU: exit code 0, a falling cost,
the closed forms of tanh's derivatives."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(script, *args, timeout=300):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script)] + list(args), capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.mark.gpu
def test_example_mlp_trains():
    for extra in ((), ("--device-weights",)):
        out = run("mlp_synthetic_digits.py", "--steps", "151", "--batch", "100", *extra)
        costs = [float(m) for m in re.findall(r"cost ([0-9.]+)", out)]
        assert len(costs) >= 3 and costs[-1] < 0.5 * costs[0], out
        acc = float(re.search(r"accuracy on 2000 fresh samples: ([0-9.]+)", out).group(1))
        assert acc > 90.0, out


@pytest.mark.gpu
def test_example_char_rnn_learns_and_samples(tmp_path):
    corpus = tmp_path / "corpus.txt"
    corpus.write_text("to be, or not to be, that is the question: whether 'tis nobler in the mind to suffer. " * 30)
    for extra in ((), ("--corpus", str(corpus))):
        out = run("char_rnn_textgen.py", "--steps", "81", "--hidden", "64", "--unroll", "12", *extra)
        costs = [float(m) for m in re.findall(r"cost ([0-9.]+)", out)]
        assert len(costs) == 3 and costs[-1] < costs[0], out
        assert len(re.findall(r"\n    \S", out)) >= 3, out          # a sampled line under every report


@pytest.mark.gpu
def test_example_higher_order_tanh():
    out = run("higher_order_tanh.py")
    assert len(re.findall(r"order \d: max \|graph - closed form\|", out)) == 5, out
