# modules that exist here win; the rest of this package resolves from a reference checkout (see src/__init__.py)
from src import extend_path

__path__ = extend_path(__path__, __name__)
