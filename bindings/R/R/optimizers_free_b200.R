## optimizers_free_b200.R - free-mode optimizers of the stochQN R package over the B200 library.
##
## Drop-in replacements for the parts of the reference's R layer that touch memory:
##   * R/allocators.R (create.r.oLBFGS / create.r.SQN / create.r.adaQN: R vectors for s_mem, y_mem, ... )  ->  one
##     external pointer to a GPU workspace per optimizer (create.b200.*)
##   * the .Call sites of run_oLBFGS_free / run_SQN_free / run_adaQN_free (R/optimizers_free.R:415-623)      ->  the
##     same three functions below, same arguments, same returned list(task, requested_on, info)
## Everything else of the reference's R layer (constructors oLBFGS_free / SQN_free / adaQN_free with their argument
## checks, update_gradient / update_hess_vec / update_fun, the guided classes of R/optimizers_guided.R and
## R/logistic.R, helpers get.task / get.iter.info / get.x.changed / check.x.and.step.size of R/helpers.R) is host
## orchestration and runs unchanged on top of these.
##
## Not runnable in this repository's image (no R); written against src/Rwrapper_b200.c.

create.b200.oLBFGS <- function(n, mem_size, hess_init, y_reg, min_curvature, check_nan, nthreads) {
	ws <- .Call("r_b200_init_oLBFGS", as.integer(n), as.integer(mem_size), as.numeric(hess_init), as.numeric(y_reg),
				as.numeric(min_curvature), as.integer(as.logical(check_nan)), as.integer(nthreads))
	list(ws = ws, n = as.integer(n), niter = integer(1), section = integer(1),
		 BFGS_mem = list(mem_size = as.integer(mem_size), mem_used = integer(1), mem_st_ix = integer(1)))
}

create.b200.SQN <- function(n, mem_size, bfgs_upd_freq, min_curvature, use_grad_diff, y_reg, check_nan, nthreads) {
	ws <- .Call("r_b200_init_SQN", as.integer(n), as.integer(mem_size), as.integer(bfgs_upd_freq), as.numeric(min_curvature),
				as.integer(as.logical(use_grad_diff)), as.numeric(y_reg), as.integer(as.logical(check_nan)), as.integer(nthreads))
	list(ws = ws, n = as.integer(n), niter = integer(1), section = integer(1), use_grad_diff = as.logical(use_grad_diff),
		 BFGS_mem = list(mem_size = as.integer(mem_size), mem_used = integer(1), mem_st_ix = integer(1),
						 upd_freq = as.integer(bfgs_upd_freq)))
}

create.b200.adaQN <- function(n, mem_size, fisher_size, bfgs_upd_freq, max_incr, min_curvature, scal_reg, rmsprop_weight,
							  use_grad_diff, y_reg, check_nan, nthreads) {
	ws <- .Call("r_b200_init_adaQN", as.integer(n), as.integer(mem_size), as.integer(fisher_size), as.integer(bfgs_upd_freq),
				as.numeric(max_incr), as.numeric(min_curvature), as.numeric(scal_reg), as.numeric(rmsprop_weight),
				as.integer(as.logical(use_grad_diff)), as.numeric(y_reg), as.integer(as.logical(check_nan)), as.integer(nthreads))
	list(ws = ws, n = as.integer(n), niter = integer(1), section = integer(1), f_prev = numeric(1),
		 use_grad_diff = as.logical(use_grad_diff),
		 BFGS_mem = list(mem_size = as.integer(mem_size), mem_used = integer(1), mem_st_ix = integer(1),
						 upd_freq = as.integer(bfgs_upd_freq)),
		 Fisher_mem = list(mem_size = as.integer(fisher_size), mem_used = integer(1), mem_st_ix = integer(1)))
}

first.request <- function(x) {
	list(task = get.task(101), requested_on = x,
		 info = list(x_changed_in_run = get.x.changed(0), iteration_number = 0, iteration_info = get.iter.info(200)))
}

run_oLBFGS_free <- function(optimizer, x, step_size) {
	check.x.and.step.size(x, step_size)
	if (!("oLBFGS_free" %in% class(optimizer))) stop("This function only applies to free-mode oLBFGS optimizer.")
	if (!optimizer$initialized) {
		p <- optimizer$saved_params
		obj <- create.b200.oLBFGS(NROW(x), p$mem_size, p$hess_init, p$y_reg, p$min_curvature, p$check_nan, p$nthreads)
		grad_init <- vector(mode = "numeric", length = obj$n)
		eval.parent(substitute(optimizer[["oLBFGS"]] <- obj))
		eval.parent(substitute(optimizer[["initialized"]] <- TRUE))
		eval.parent(substitute(optimizer[["saved_params"]] <- NULL))
		eval.parent(substitute(optimizer[["gradient"]] <- grad_init))
		return(first.request(x))
	}
	o <- optimizer$oLBFGS
	if (NROW(x) != o$n) stop("'x' has wrong dimensions.")
	req <- vector(mode = "numeric", length = o$n)
	x_changed <- integer(1); task <- integer(1); iter_info <- integer(1)
	## x and optimizer$gradient are modified in place, as in the reference
	.Call("r_b200_run_oLBFGS", o$ws, x, optimizer$gradient, as.numeric(step_size),
		  o$niter, o$section, o$BFGS_mem$mem_used, o$BFGS_mem$mem_st_ix, x_changed, req, task, iter_info)
	list(task = get.task(task), requested_on = req,
		 info = list(x_changed_in_run = get.x.changed(x_changed), iteration_number = as.integer(o$niter) + 1 - 1,
					 iteration_info = get.iter.info(iter_info)))
}

run_SQN_free <- function(optimizer, x, step_size) {
	check.x.and.step.size(x, step_size)
	if (!("SQN_free" %in% class(optimizer))) stop("This function only applies to free-mode SQN optimizer.")
	if (!optimizer$initialized) {
		p <- optimizer$saved_params
		obj <- create.b200.SQN(NROW(x), p$mem_size, p$bfgs_upd_freq, p$min_curvature, p$use_grad_diff, p$y_reg,
							   p$check_nan, p$nthreads)
		grad_init <- vector(mode = "numeric", length = obj$n)
		hv_init <- vector(mode = "numeric", length = obj$n)
		eval.parent(substitute(optimizer[["SQN"]] <- obj))
		eval.parent(substitute(optimizer[["initialized"]] <- TRUE))
		eval.parent(substitute(optimizer[["saved_params"]] <- NULL))
		eval.parent(substitute(optimizer[["gradient"]] <- grad_init))
		eval.parent(substitute(optimizer[["hess_vec"]] <- hv_init))
		return(first.request(x))
	}
	o <- optimizer$SQN
	if (NROW(x) != o$n) stop("'x' has wrong dimensions.")
	req <- vector(mode = "numeric", length = o$n)
	req_vec <- vector(mode = "numeric", length = o$n)
	x_changed <- integer(1); task <- integer(1); iter_info <- integer(1)
	.Call("r_b200_run_SQN", o$ws, x, optimizer$gradient, optimizer$hess_vec, as.numeric(step_size),
		  o$niter, o$section, o$BFGS_mem$mem_used, o$BFGS_mem$mem_st_ix, x_changed, req, req_vec, task, iter_info)
	requested_on <- if (task == 104) list(req, req_vec) else req
	list(task = get.task(task), requested_on = requested_on,
		 info = list(x_changed_in_run = get.x.changed(x_changed), iteration_number = as.integer(o$niter) + 1 - 1,
					 iteration_info = get.iter.info(iter_info)))
}

run_adaQN_free <- function(optimizer, x, step_size) {
	check.x.and.step.size(x, step_size)
	if (!("adaQN_free" %in% class(optimizer))) stop("This function only applies to free-mode adaQN optimizer.")
	if (!optimizer$initialized) {
		p <- optimizer$saved_params
		obj <- create.b200.adaQN(NROW(x), p$mem_size, p$fisher_size, p$bfgs_upd_freq, p$max_incr, p$min_curvature,
								 p$scal_reg, p$rmsprop_weight, p$use_grad_diff, p$y_reg, p$check_nan, p$nthreads)
		grad_init <- vector(mode = "numeric", length = obj$n)
		eval.parent(substitute(optimizer[["adaQN"]] <- obj))
		eval.parent(substitute(optimizer[["initialized"]] <- TRUE))
		eval.parent(substitute(optimizer[["saved_params"]] <- NULL))
		eval.parent(substitute(optimizer[["gradient"]] <- grad_init))
		eval.parent(substitute(optimizer[["f"]] <- numeric(1)))
		return(first.request(x))
	}
	o <- optimizer$adaQN
	if (NROW(x) != o$n) stop("'x' has wrong dimensions.")
	req <- vector(mode = "numeric", length = o$n)
	x_changed <- integer(1); task <- integer(1); iter_info <- integer(1)
	.Call("r_b200_run_adaQN", o$ws, x, as.numeric(optimizer$f), optimizer$gradient, as.numeric(step_size),
		  o$niter, o$section, o$BFGS_mem$mem_used, o$BFGS_mem$mem_st_ix,
		  o$Fisher_mem$mem_used, o$Fisher_mem$mem_st_ix, o$f_prev, x_changed, req, task, iter_info)
	list(task = get.task(task), requested_on = req,
		 info = list(x_changed_in_run = get.x.changed(x_changed), iteration_number = as.integer(o$niter) + 1 - 1,
					 iteration_info = get.iter.info(iter_info)))
}
