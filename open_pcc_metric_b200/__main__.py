from . import handler

if __name__ == "__main__":
    handler.cli()  # pylint: disable=no-value-for-parameter
