from FidelityFusion_Models.AR_autoRegression import AR, train_AR
from FidelityFusion_Models.CIGAR import CIGAR, train_CIGAR
from FidelityFusion_Models.CAR_ContinuousAutoRegression import ContinuousAutoRegression, train_CAR
# from CAR_ContinuousAutoRegression_Large import ContinuousAutoRegression, train_CAR
from FidelityFusion_Models.GAR import GAR, train_GAR
from FidelityFusion_Models.MF_data import MultiFidelityDataManager
from FidelityFusion_Models.NAR import NAR, train_NAR
from FidelityFusion_Models.ResGP import ResGP, train_ResGP
