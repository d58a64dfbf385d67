from . import seed_everything  # noqa: F401
