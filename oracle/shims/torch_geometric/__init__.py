"""Minimal stand-in for torch-geometric==1.7.0 (requirements.txt:9 of the reference): only the
symbols the validate path imports.  TEST INFRASTRUCTURE for oracle/gen_golden.py only."""
from . import data, transforms, utils, nn, typing  # noqa
