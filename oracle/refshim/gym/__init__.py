"""Stand-in for gym 0.26 (absent in this image): only what CybORG CC4 touches at import / space construction.
Test infrastructure for running the UNMODIFIED reference from /root/reference (see oracle/README.md)."""
from . import spaces, utils, vector
from .spaces import Space
