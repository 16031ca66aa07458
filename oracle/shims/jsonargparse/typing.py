Path_fr = str
Path_dw = str
Path_fc = str
