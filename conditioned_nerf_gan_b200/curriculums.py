"""Curriculum plumbing of the train step: which hyper-parameters are in force at a given step.

Mirror of the helper functions of ``configs/curriculums.py:83-151`` (the schema is documented there: integer keys are
upsample steps whose dicts override the running settings, every other key is a global setting).  The configs themselves
(``configs/thousand/*.py``) are data and stay with the caller; these functions only read them.
"""
from __future__ import annotations

from typing import Dict


def _steps(curriculum: Dict):
    return sorted(k for k in curriculum.keys() if type(k) == int)


def extract_metadata(curriculum: Dict, current_step: int) -> Dict:
    """Settings of the latest upsample step <= ``current_step`` plus all non-integer keys (curriculums.py:124-137)."""
    out: Dict = {}
    for step in reversed(_steps(curriculum)):
        if step <= current_step:
            out.update(curriculum[step])
            break
    out.update({k: v for k, v in curriculum.items() if type(k) != int})
    return out


def next_upsample_step(curriculum: Dict, current_step: int):
    """First later step whose image size exceeds the current one, else +inf (curriculums.py:83-94)."""
    size = extract_metadata(curriculum, current_step)["img_size"]
    for step in _steps(curriculum):
        if step > current_step and curriculum[step].get("img_size", 512) > size:
            return step
    return float("Inf")


def last_upsample_step(curriculum: Dict, current_step: int) -> int:
    """First step <= ``current_step`` that already had the current image size (curriculums.py:97-108)."""
    size = extract_metadata(curriculum, current_step)["img_size"]
    for step in _steps(curriculum):
        if step <= current_step and curriculum[step]["img_size"] == size:
            return step
    return 0


def update_recursive(dict1: Dict, dict2: Dict) -> Dict:
    """Overlay ``dict2`` on ``dict1`` in place, descending into nested dicts (curriculums.py:140-155)."""
    for k, v in dict2.items():
        if k not in dict1:
            dict1[k] = dict()
        if isinstance(v, dict):
            update_recursive(dict1[k], v)
        else:
            dict1[k] = v
    return dict1
