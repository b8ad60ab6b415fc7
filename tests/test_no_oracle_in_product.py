"""CPU: the product package never imports, calls or shells out to oracle/ (test infrastructure only)."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_package_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "traversome_b200")
    offenders = []
    for dirpath, _dirs, files in os.walk(pkg):
        for name in files:
            path = os.path.join(dirpath, name)
            if name.endswith(".py"):
                tree = ast.parse(open(path).read())
                for node in ast.walk(tree):
                    mods = []
                    if isinstance(node, ast.Import):
                        mods = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        mods = [node.module or ""]
                    if any(m == "oracle" or m.startswith("oracle.") for m in mods):
                        offenders.append(path)
            elif name.endswith((".cu", ".cuh", ".h", ".sh")):
                if "oracle" in open(path).read():
                    offenders.append(path)
    assert not offenders, offenders


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from traversome_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "libtvs.so"))
    try:
        _cabi.load()
    except ImportError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("load() must raise when libtvs.so is missing")
