from .group import World, get_world, init_world, shutdown, split_clients  # noqa: F401
