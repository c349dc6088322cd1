from .single_snake import SingleSnake  # noqa: F401
from .multi_snake import MultiSnake  # noqa: F401
from .simple_gridworld import SimpleGridworld  # noqa: F401
