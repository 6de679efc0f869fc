from v1.CFKG import discrete_fidelity_knowledgement_gradient
from v1.MF_EI import expected_improvement
from v1.MF_ES import entropy_search
from v1.MF_UCB import upper_confidence_bound
