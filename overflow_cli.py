"""Drop-in for the reference's src/overflow_cli.py (D8 commands only)."""
from overflow_b200.cli import flow_accumulation_cli, flow_direction_cli, flow_routing_cli, main  # noqa: F401

if __name__ == "__main__":
    main()  # pylint: disable=no-value-for-parameter
