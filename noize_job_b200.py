"""Import shim: the package directory is `noize-job_b200/` (hyphen, as the task names it), which Python
cannot import by name.  `import noize_job_b200` loads that directory as a regular package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "noize-job_b200")
_spec = importlib.util.spec_from_file_location("noize_job_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["noize_job_b200"] = _mod
_spec.loader.exec_module(_mod)
