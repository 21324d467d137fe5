from pytorch_probabilisticlayers_b200.data import (generate_linreg_data,  # noqa: F401
                                                  generate_nonstationary_data)
