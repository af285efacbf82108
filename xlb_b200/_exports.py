"""Namespace helper: the sub-packages of xlb_b200 re-export their public classes under the same names the reference
package uses (`xlb.operator.boundary_condition.EquilibriumBC`, ...).  `export` imports the listed sibling modules in
order and copies the listed names into the package namespace."""

from importlib import import_module


def export(namespace: dict, package: str, table: dict) -> list:
    names = []
    for module, symbols in table.items():
        mod = import_module(f"{package}.{module}")
        for sym in symbols:
            namespace[sym] = getattr(mod, sym)
            names.append(sym)
    namespace["__all__"] = names
    return names
