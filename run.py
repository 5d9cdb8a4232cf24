"""Command-line entry point with the reference's flags: python run.py -m <models> [-cf] [-cpu] ..."""
from innfer_b200.run import main

if __name__ == "__main__":
    main()
