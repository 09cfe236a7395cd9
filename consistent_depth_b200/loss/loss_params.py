"""Mirror of the reference's loss/loss_params.py:5-40 (CLI flags + run-tag string)."""


class LossParams:
    @staticmethod
    def add_arguments(parser):
        parser.add_argument("--lambda_view_baseline", type=float, default=-1,
                            help="weight of the disparity term; < 0 selects the model adapter's default")
        parser.add_argument("--lambda_reprojection", type=float, default=1.0,
                            help="weight of the reprojection term")
        parser.add_argument("--lambda_parameter", type=float, default=0,
                            help="weight of the L1 parameter-drift regulariser")
        return parser

    @staticmethod
    def make_str(opt):
        return f"B{opt.lambda_view_baseline}_R{opt.lambda_reprojection}_PL1-{opt.lambda_parameter}"
