from .single_snake import SingleSnake  # noqa: F401
