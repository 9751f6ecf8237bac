"""Plugin registry with the reference's plugin names (the drop-in surface).

Reference registration sites: transientpath.py:329, transientnlospath.py:930-931,
transient_prbvolpath.py:534-535, transient_hdr_film.py:314, nloscapturemeter.py:245.
"""
from __future__ import annotations

from typing import Any, Callable, Dict

from .scene import Properties

INTEGRATORS: Dict[str, Callable[[Properties], Any]] = {}
FILMS: Dict[str, Callable[[Properties], Any]] = {}
SENSORS: Dict[str, Callable[[Properties], Any]] = {}


def register_integrator(name: str, factory: Callable[[Properties], Any]) -> None:
    INTEGRATORS[name] = factory


def register_film(name: str, factory: Callable[[Properties], Any]) -> None:
    FILMS[name] = factory


def register_sensor(name: str, factory: Callable[[Properties], Any]) -> None:
    SENSORS[name] = factory


def instantiate(d: Dict[str, Any]):
    t = d.get("type")
    props = Properties({k: v for k, v in d.items() if k != "type"})
    for table in (INTEGRATORS, FILMS, SENSORS):
        if t in table:
            return table[t](props)
    raise ValueError(f"unknown plugin type {t!r}; registered: "
                     f"{sorted(INTEGRATORS) + sorted(FILMS) + sorted(SENSORS)}")
