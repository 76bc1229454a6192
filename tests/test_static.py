"""Branches that only run on a GPU box (multi-rank paths, the detect arm, error messages) are never executed by the CPU
suite: at least every GLOBAL name they load must exist.  Walks the code objects of bench.py, __graft_entry__.py and every
module of the package and checks LOAD_GLOBAL targets against the module namespace and the builtins."""
import builtins
import dis
import importlib
import importlib.util
import os
import pkgutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _code_objects(code):
    yield code
    for c in code.co_consts:
        if isinstance(c, types.CodeType):
            yield from _code_objects(c)


def _missing_globals(module, source_path):
    src = open(source_path).read()
    top = compile(src, source_path, "exec")
    known = set(vars(module)) | set(dir(builtins)) | {"__class__"}
    missing = []
    for code in _code_objects(top):
        for ins in dis.get_instructions(code):
            # LOAD_NAME (module and class bodies) is resolved while importing, and the import succeeded
            if ins.opname == "LOAD_GLOBAL" and isinstance(ins.argval, str) and ins.argval not in known:
                missing.append((code.co_name, ins.argval, ins.positions.lineno if ins.positions else None))
    return missing


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules.setdefault(name, mod)
    spec.loader.exec_module(mod)
    return mod


def test_every_global_name_used_in_bench_and_the_package_exists():
    sys.path.insert(0, ROOT)
    problems = {}
    for fname, name in (("bench.py", "_static_bench"), ("__graft_entry__.py", "_static_graft_entry")):
        path = os.path.join(ROOT, fname)
        miss = _missing_globals(_load(path, name), path)
        if miss:
            problems[fname] = miss
    import sos_wsod_b200

    for info in pkgutil.walk_packages(sos_wsod_b200.__path__, "sos_wsod_b200."):
        mod = importlib.import_module(info.name)
        path = getattr(mod, "__file__", None)
        if path and path.endswith(".py"):
            miss = _missing_globals(mod, path)
            if miss:
                problems[info.name] = miss
    assert not problems, problems
