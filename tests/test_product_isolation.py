"""The oracle is test infrastructure: nothing the product package ships may import, link or execute it (nor the g++ emulation
harness under tests/emul, nor the reference tree).  Checked statically over every source of the package and dynamically in a fresh
interpreter; a missing libgracing.so must be an ImportError, never a CPU route."""
import ast
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "generalizableracing_b200")
FORBIDDEN_ROOTS = {"oracle", "tests"}


def _sources(suffixes):
    for base, dirs, files in os.walk(PKG):
        dirs[:] = [d for d in dirs if d not in ("build", "__pycache__")]
        for f in files:
            if f.endswith(suffixes):
                yield os.path.join(base, f)


def test_no_python_source_of_the_package_imports_the_oracle():
    seen = 0
    for path in _sources((".py",)):
        tree = ast.parse(open(path).read(), path)
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom) and node.level == 0 and node.module:
                names = [node.module]
            for n in names:
                assert n.split(".")[0] not in FORBIDDEN_ROOTS, (path, n)
            # importlib.import_module("oracle...") / __import__("oracle...")
            if isinstance(node, ast.Call) and node.args and isinstance(node.args[0], ast.Constant) and isinstance(node.args[0].value, str):
                fn = node.func.attr if isinstance(node.func, ast.Attribute) else getattr(node.func, "id", "")
                if fn in ("import_module", "__import__"):
                    assert node.args[0].value.split(".")[0] not in FORBIDDEN_ROOTS, (path, node.args[0].value)
        seen += 1
    assert seen >= 10


def test_no_native_source_of_the_package_includes_the_oracle_or_the_reference():
    seen = 0
    for path in _sources((".cu", ".cuh", ".h", ".cpp")):
        for i, line in enumerate(open(path), 1):
            s = line.strip()
            if s.startswith("#include"):
                assert "oracle" not in s and "/root/reference" not in s and "tests/" not in s, (path, i, s)
        seen += 1
    assert seen >= 10


def test_importing_the_whole_package_loads_no_oracle_module():
    code = (
        "import sys, pkgutil, importlib, generalizableracing_b200 as p\n"
        "for m in pkgutil.walk_packages(p.__path__, p.__name__ + '.'):\n"
        "    if '.build' in m.name or m.name.endswith('libgracing'): continue   # the C-ABI library is not a Python module\n"
        "    importlib.import_module(m.name)\n"
        "bad = [k for k in sys.modules if k.split('.')[0] in ('oracle', 'tests')]\n"
        "assert not bad, bad\n"
        "print('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_missing_library_is_an_import_error_not_a_cpu_route():
    code = (
        "import generalizableracing_b200._lib as L\n"
        "L.LIB_PATH = L.LIB_PATH + '.absent'\n"
        "try:\n"
        "    L.load()\n"
        "except ImportError as e:\n"
        "    assert 'no CPU fallback' in str(e); print('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
