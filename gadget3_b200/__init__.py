"""gadget3_b200: B200-native batched objective evaluator for gadget3 models.

Host-side mirror of the gadget3 model-definition API (g3_stock, g3a_*, g3l_*, g3_parameterized,
g3_init_val) plus the new backend ``g3_to_cuda()`` / ``g3_cuda_adfun()`` that sits beside the
reference's ``g3_to_r()`` / ``g3_to_tmb()`` (R/run_r.R, R/run_tmb.R). The compute path is the
hand-written CUDA library in ``csrc/`` reached through the C ABI in ``include/g3cuda.h``.
"""
from .formula import (G3Unsupported, ParameterTemplate, Lookup, age, area, cur_step, cur_step_size,
                      cur_year, g3_as_integer, g3_avoid_zero, g3_exp, g3_init_val, g3_log, g3_parameterized, g3_sqrt,
                      g3_timevariable)
from .stock import (g3_areas, g3_fleet, g3_stock, g3_stock_def, g3s_age, g3s_length, g3s_livesonareas)
from .actions import (g3_action_order, g3_suitability_andersen, g3_suitability_andersenfleet,
                      g3_suitability_constant, g3_suitability_exponential, g3_suitability_exponentiall50,
                      g3_suitability_gamma, g3_suitability_richards, g3_suitability_straightline, g3_timeareadata, g3a_age, g3a_grow_impl_bbinom,
                      g3a_grow_lengthvbsimple, g3a_grow_weightsimple, g3a_growmature,
                      g3a_initialconditions_normalcv, g3a_initialconditions_normalparam,
                      g3a_mature_constant, g3a_mature_continuous, g3a_migrate, g3a_naturalmortality, g3a_otherfood, g3a_spawn,
                      g3a_spawn_recruitment_bevertonholt, g3a_spawn_recruitment_bevertonholt_ss3, g3a_spawn_recruitment_fecundity,
                      g3a_spawn_recruitment_hockeystick, g3a_spawn_recruitment_ricker, g3a_spawn_recruitment_simplessb,
                      g3a_naturalmortality_exp, g3a_predate, g3a_predate_catchability_predator,
                      g3a_predate_catchability_effortfleet, g3a_predate_catchability_linearfleet,
                      g3a_predate_catchability_numberfleet, g3a_predate_catchability_totalfleet, g3a_predate_fleet,
                      g3a_predate_maxconsumption, g3a_renewal_initabund, g3a_renewal_normalcv,
                      g3a_renewal_normalparam, g3a_renewal_vonb, g3a_renewal_vonb_recl,
                      g3a_renewal_vonb_t0, g3a_report_detail, g3a_time)
from .likelihood import (g3l_abundancedistribution, g3l_bounds_penalty, g3l_catchdistribution,
                         g3l_distribution, g3l_distribution_multinomial, g3l_distribution_sumofsquaredlogratios, g3l_distribution_sumofsquares,
                         g3l_distribution_surveyindices_linear, g3l_distribution_surveyindices_log,
                         g3l_random_dnorm, g3l_random_walk, g3l_understocking)
from .run_cuda import G3CudaADFun, G3CudaModel, g3_cuda_adfun, g3_to_cuda, unpack_report

__version__ = "0.1.0"
