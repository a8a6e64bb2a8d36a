"""Runge-Kutta tableaux: containers (``common``), parametric families and the named methods (``providers``).

Re-exports the names of skrample/sampling/tableaux/__init__.py so ``tableaux.RK2.Mid`` etc. resolve as in the reference.
"""

from collections.abc import Sequence

from . import providers as _named
from .common import EmbeddedTableau, Tableau, TableauType

TableauProvider = _named.TableauProvider
CustomTableau = _named.CustomTableau
RK2Custom, RK3Custom, RK4Custom = _named.RK2Custom, _named.RK3Custom, _named.RK4Custom
RK1, RK2, RK3, RK4 = _named.RK1, _named.RK2, _named.RK3, _named.RK4
RKE2, RKE3, RKE5 = _named.RKE2, _named.RKE3, _named.RKE5
RKZ, SSP, WSO, Shanks1965 = _named.RKZ, _named.SSP, _named.WSO, _named.Shanks1965

BUILTIN_TABLEAUX: Sequence[TableauProvider[Tableau]] = [method for family in (RK1, RK2, RK3, RK4, RKZ, SSP) for method in family]
"Every usable explicit method"
BUILTIN_EMBEDDED_TABLEAU: Sequence[TableauProvider[EmbeddedTableau]] = [pair for family in (RKE2, RKE3, RKE5) for pair in family]
"Every usable embedded pair"
GRAVEYARD: Sequence[TableauProvider[TableauType]] = [method for family in (WSO, Shanks1965) for method in family]
"Methods kept for completeness that sample poorly"

__all__ = [
    "BUILTIN_EMBEDDED_TABLEAU", "BUILTIN_TABLEAUX", "GRAVEYARD", "RK1", "RK2", "RK3", "RK4", "RKE2", "RKE3", "RKE5", "RKZ", "SSP", "WSO",
    "CustomTableau", "EmbeddedTableau", "RK2Custom", "RK3Custom", "RK4Custom", "Shanks1965", "Tableau", "TableauProvider", "TableauType",
]  # fmt: skip
