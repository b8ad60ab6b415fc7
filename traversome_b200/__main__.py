"""``python -m traversome_b200 thorough ...`` = the reference's ``traversome thorough ...`` with the ML-fit path on the GPU."""
from .dropin import main

if __name__ == "__main__":
    main()
