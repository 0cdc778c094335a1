"""Stand-in for the one `sbi` symbol the reference toy tasks import (BoxUniform). Test infra only."""
