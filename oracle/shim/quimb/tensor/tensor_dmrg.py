class DMRG2:
    pass
