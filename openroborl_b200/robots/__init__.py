"""Per-robot constant modules, same names as the reference's robot modules."""
from . import laikago, mini_cheetah  # noqa: F401


def get(name_robot):
    """Minitaur.__init__ robot selection (reference minitaur.py:92-97).

    Also accepts "minicheetah": run.py:57 derives that spelling from the task name and the reference
    then fails with ValueError("wrong robot select").
    """
    if name_robot == "laikago":
        return laikago
    if name_robot in ("mini_cheetah", "minicheetah"):
        return mini_cheetah
    raise ValueError("wrong robot select")
