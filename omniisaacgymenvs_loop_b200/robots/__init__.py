"""Stand-ins for the simulator-side objects the USV task talks to (Isaac Sim articulation views)."""
from .heron_view import PlanarHeronView, PlanarWorld  # noqa: F401
