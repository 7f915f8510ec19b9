"""The example scripts need a B200 to run; here they are only parsed, and every name they take from the package is checked to exist
(imports `from tnb200... import x`, attributes `tnb200.x` / `models.x`)."""
import ast
import importlib
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = sorted(f for f in os.listdir(os.path.join(ROOT, "examples")) if f.endswith(".py"))


@pytest.mark.parametrize("name", EXAMPLES)
def test_example_uses_only_existing_names(name):
    import tnb200
    tree = ast.parse(open(os.path.join(ROOT, "examples", name)).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("tnb200"):
            mod = importlib.import_module(node.module)
            for a in node.names:
                assert hasattr(mod, a.name), f"{name}: {node.module}.{a.name} does not exist"
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id in ("tnb200", "models"):
            mod = tnb200 if node.value.id == "tnb200" else tnb200.models
            assert hasattr(mod, node.attr), f"{name}: {node.value.id}.{node.attr} does not exist"
