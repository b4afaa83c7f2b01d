"""bench.py and the GPU job helpers cannot run without a device, so a typo in a rarely taken branch (the line is assembled AFTER the
timed region) would only show on the GPU box.  Static check: every name a function loads is bound somewhere it can see."""
import ast
import builtins
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bound_names(node):
    out = set()
    for n in ast.walk(node):
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            out.add(n.id)
        elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            out.add(n.name)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                out.add((a.asname or a.name).split('.')[0])
        elif isinstance(n, ast.arg):
            out.add(n.arg)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            out.add(n.name)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            out.update(n.names)
    return out


def unbound_loads(path):
    tree = ast.parse(open(path).read())
    module_names = bound_names(tree) | set(dir(builtins)) | {'__file__', '__name__', '__doc__'}
    bad = []
    for fn in ast.walk(tree):
        if isinstance(fn, (ast.FunctionDef, ast.AsyncFunctionDef)):
            local = bound_names(fn)
            for n in ast.walk(fn):
                if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in local and n.id not in module_names:
                    bad.append('%s:%d %s' % (os.path.basename(path), n.lineno, n.id))
    for n in ast.walk(tree):
        if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in module_names:
            bad.append('%s:%d %s' % (os.path.basename(path), n.lineno, n.id))
    return sorted(set(bad))


@pytest.mark.parametrize('rel', ['bench.py', '__graft_entry__.py', 'tools/one_pass.py', 'tools/gemm_k_sweep.py', 'tools/ozaki_gpu_check.py',
                                 'tools/summarize_launches.py', 'tools/sass_mix.py', 'tests/multi_worker.py',
                                 'bundle-adjustment_b200/verify.py', 'bundle-adjustment_b200/_lib.py', 'bundle-adjustment_b200/workloads.py'])
def test_no_unbound_names(rel):
    assert unbound_loads(os.path.join(ROOT, rel)) == []


def test_bench_line_carries_the_contract_keys():
    """The keys of the measurement contract are all spelled out in bench.py's result line (both arms)."""
    src = open(os.path.join(ROOT, 'bench.py')).read()
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype',
                'data', 'config', 'workload', 'roofline', 'bound', 'achieved', 'peak', 'frac', 'traffic', 'cpu_baseline', 'cores', 'kind',
                'sample', 'e2e', 'h2d_bytes_per_step', 'd2h_bytes_per_step', 'clocks', 'sm_mhz', 'sm_max_mhz', 'reasons', 'gpu_launches',
                'impl', 'check'):
        assert "'%s'" % key in src, key


def test_oracle_and_reference_stay_out_of_the_product():
    """oracle/ is test infrastructure: the package, the tools and the C / C++ / CUDA sources never import, link or read it; bench.py only
    inside its CPU-baseline / reference-arm functions; and nothing that runs on the GPU box reads /root/reference (only the
    fixture generators under tests/golden/ do, in the build container)."""
    import glob
    import re
    pkg = glob.glob(os.path.join(ROOT, 'bundle-adjustment_b200', '**', '*'), recursive=True)
    for path in pkg + glob.glob(os.path.join(ROOT, 'tools', '*')) + glob.glob(os.path.join(ROOT, 'include', '*')):
        if os.path.isfile(path) and path.endswith(('.py', '.cu', '.cuh', '.h', '.hpp', '.cpp', '.sh')):
            src = open(path).read()
            assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), path
            assert 'libjaicov_oracle' not in src and 'jaicov_oracle.c' not in src, path
            assert not re.search(r'(open|listdir|exists|join|load|glob|walk)\([^\n]*/root/reference', src), path      # citations in comments are fine
    tree = ast.parse(open(os.path.join(ROOT, 'bench.py')).read())
    importers = set()
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for n in ast.walk(fn):
            if isinstance(n, ast.ImportFrom) and (n.module or '').split('.')[0] == 'oracle':
                importers.add(fn.name)
    assert importers and importers <= {'cpu_sample', 'best_effort_cpu', 'reference_arm'}, importers
    for n in tree.body:                                   # no module-level import of the oracle
        assert not (isinstance(n, (ast.Import, ast.ImportFrom)) and 'oracle' in ast.dump(n))
    for path in [os.path.join(ROOT, 'bench.py'), os.path.join(ROOT, '__graft_entry__.py')] + glob.glob(os.path.join(ROOT, 'tests', '*.py')):
        if os.path.basename(path) == 'test_bench_static.py':
            continue
        src = open(path).read()
        uses = [l for l in src.splitlines() if re.search(r'(open|listdir|exists|join|load|glob|walk)\([^\n]*/root/reference', l)]
        assert not uses, (path, uses)
