"""Drop-in shim: put ``<repo>/dropin`` (and the repo root) on PYTHONPATH and existing
``import cbioseq`` code picks up the B200-native extension instead of the reference's."""
from bioseq_b200.cbioseq import *  # noqa: F401,F403
from bioseq_b200.cbioseq import Tokenizer, Threading, set_num_threads, get_num_threads  # noqa: F401
